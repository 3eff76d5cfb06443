"""TEST INFRASTRUCTURE — CPU oracle (numpy f64 + OpenCV) for the geometry stages.  Never imported by the product.

  undistort_points      sfm/geometry.py:103-118  -> cv2.undistortPoints(pts, K, dist, None, K), cast to f32.
                        OpenCV (third party, unpinned in pyproject.toml:22; 4.13.0 here) is the reference's own
                        arithmetic for this call, so the oracle calls it directly; `undistort_points_np` restates
                        OpenCV's published algorithm (5 fixed-point iterations of the Brown model) for the
                        stopping rule the CUDA kernel mirrors.
  iterative_ls          thirdparty/triangulation.py:79-177 (vectorised over points; 4x3 LS solve per iteration;
                        NOTE the reference re-scales A and b *cumulatively* every iteration, :157-160)
  linear_dlt            sfm/triangulation.py:154-183 (6x6 SVD null vector, dehomogenised)
  fundamental_magsac    matching/geometric_verification.py:89-92 -> cv2.findFundamentalMat(USAC_MAGSAC, 0.5, 0.999, 100000)
  magsac_weights / magsac_quality / magsac_polish
                        MAGSAC++ (Barath et al. 2020) as OpenCV's USAC applies it to F (opencv/modules/calib3d/src/usac/
                        quality.cpp; third party, not under /root/reference): sigma-consensus++ weights and marginalised
                        quality on the squared Sampson error (4 DoF, k = 3.64), and the re-weighted normalised 8-point
                        polisher iterated to its fixed point.  cv2 does not expose sigma_max; MAGSAC_CUTOFF_PX = 4.5 is
                        fitted to cv2's own output (tests/test_oracle_vs_golden.py::test_magsac_polisher_pinned_to_opencv,
                        scripts/magsac_probe.py).
  helmert_linear / apply_transform
                        sfm/absolute_orientation.py:141-154,247-287 + thirdparty/transformations.py:889-1020 (Horn's quaternion
                        method, affine_matrix_from_points with shear=False, usesvd=False)
  sampson_error / symmetric epipolar distance: used by tests to compare inlier sets geometrically.
Pinned against the reference itself by tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import cv2
import numpy as np


def undistort_points(pts: np.ndarray, K: np.ndarray, dist: np.ndarray) -> np.ndarray:
    return cv2.undistortPoints(pts, K, dist, None, K)[:, 0, :].astype("float32")


def undistort_points_np(pts: np.ndarray, K: np.ndarray, dist: np.ndarray, iters: int = 5) -> np.ndarray:
    """OpenCV's cvUndistortPointsInternal with default criteria (5 iterations), k1,k2,p1,p2,k3 model, f64."""
    k1, k2, p1, p2, k3 = [float(v) for v in np.asarray(dist).ravel()[:5]]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    x0 = (pts[:, 0].astype(np.float64) - cx) / fx
    y0 = (pts[:, 1].astype(np.float64) - cy) / fy
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    return np.stack([x * fx + cx, y * fy + cy], 1).astype("float32")


def iterative_ls(u1: np.ndarray, P1: np.ndarray, u2: np.ndarray, P2: np.ndarray, tol: float = 3e-5):
    """Returns (X [N,3] f64, status [N] int) with the reference's status codes."""
    u1 = np.asarray(u1, dtype=np.float64)
    u2 = np.asarray(u2, dtype=np.float64)
    n = len(u1)
    X = np.zeros((n, 3))
    status = np.zeros(n, dtype=int)
    for i in range(n):
        rows, rhs = [], []
        for (u, P) in ((u1[i], P1), (u2[i], P2)):
            rows.append(u[0] * P[2, :3] - P[0, :3])
            rows.append(u[1] * P[2, :3] - P[1, :3])
            rhs.append(-(u[0] * P[2, 3] - P[0, 3]))
            rhs.append(-(u[1] * P[2, 3] - P[1, 3]))
        A, b = np.array(rows), np.array(rhs)
        d1 = d2 = 1.0
        for it in range(10):
            x = np.linalg.lstsq(A, b, rcond=None)[0]
            d1n = P1[2, :3] @ x + P1[2, 3]
            d2n = P2[2, :3] @ x + P2[2, 3]
            if abs(d1n - d1) <= tol and abs(d2n - d2) <= tol:
                break
            A[:2] /= d1n
            b[:2] /= d1n
            A[2:] /= d2n
            b[2:] /= d2n
            d1, d2 = d1n, d2n
        X[i] = x
        st = int(d1n > 0 and d2n > 0)          # `i < 10` in the reference is always true
        if d1n <= 0:
            st -= 1
        if d2n <= 0:
            st -= 2
        status[i] = st
    return X, status


def linear_dlt(x1: np.ndarray, P1: np.ndarray, x2: np.ndarray, P2: np.ndarray) -> np.ndarray:
    """x1,x2 [N,2] undistorted pixels -> X [N,3] f64."""
    out = np.zeros((len(x1), 3))
    for i in range(len(x1)):
        M = np.zeros((6, 6))
        M[:3, :4], M[3:, :4] = P1, P2
        M[:3, 4] = -np.array([x1[i, 0], x1[i, 1], 1.0])
        M[3:, 5] = -np.array([x2[i, 0], x2[i, 1], 1.0])
        V = np.linalg.svd(M)[-1]
        out[i] = V[-1, :3] / V[-1, 3]
    return out


def fundamental_magsac(m0: np.ndarray, m1: np.ndarray):
    F, inl = cv2.findFundamentalMat(m0, m1, cv2.USAC_MAGSAC, 0.5, 0.999, 100000)
    return F, (inl > 0).squeeze()


MAGSAC_K = 3.64                      # 0.99 quantile of the chi distribution with 4 degrees of freedom
MAGSAC_CUTOFF_PX = 4.5               # k * sigma_max


def _upper_gamma_1p5(x):
    from scipy.special import erfc
    sx = np.sqrt(x)
    return 0.886226925452758 * erfc(sx) + sx * np.exp(-x)


def magsac_weights(r2: np.ndarray, cutoff: float = MAGSAC_CUTOFF_PX) -> np.ndarray:
    """w(r^2) = Gamma(3/2, r^2 / (2 s^2)) - Gamma(3/2, k^2 / 2) for r < k s, else 0 (s = cutoff / k)."""
    s = cutoff / MAGSAC_K
    w = _upper_gamma_1p5(r2 / (2 * s * s)) - _upper_gamma_1p5(np.float64(MAGSAC_K ** 2 / 2))
    return np.where(r2 < cutoff * cutoff, np.maximum(w, 0.0), 0.0)


def magsac_quality(F: np.ndarray, m0: np.ndarray, m1: np.ndarray, cutoff: float = MAGSAC_CUTOFF_PX) -> float:
    """sum over r < cutoff of 1 - loss(r^2) / loss(cutoff^2),
    loss = s^2/2 * gamma(5/2, x) + r^2/4 * (Gamma(3/2, x) - Gamma(3/2, k^2/2)), x = r^2 / (2 s^2)."""
    s = cutoff / MAGSAC_K
    r2 = sampson_distance(F, m0, m1) ** 2
    r2 = r2[r2 < cutoff * cutoff]
    x = r2 / (2 * s * s)
    up15 = _upper_gamma_1p5(x)
    lo25 = 1.329340388179137 - (1.5 * up15 + x * np.sqrt(x) * np.exp(-x))
    xk = MAGSAC_K ** 2 / 2
    gk = _upper_gamma_1p5(np.float64(xk))
    max_loss = s * s / 2 * (1.329340388179137 - (1.5 * gk + xk * np.sqrt(xk) * np.exp(-xk)))
    return float(np.sum(1.0 - (s * s / 2 * lo25 + r2 / 4 * (up15 - gk)) / max_loss))


def weighted_eight_point(m0: np.ndarray, m1: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Weighted normalised 8-point fit; the similarity normalisation uses ALL correspondences (centroid, mean distance sqrt 2),
    as the CUDA path does.  Returns F with unit Frobenius norm."""
    def norm(p):
        c = p.mean(0)
        s = np.sqrt(2) / np.sqrt(((p - c) ** 2).sum(1)).mean()
        return (p - c) * s, np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1]])
    a, T0 = norm(m0.astype(np.float64))
    b, T1 = norm(m1.astype(np.float64))
    A = np.stack([b[:, 0] * a[:, 0], b[:, 0] * a[:, 1], b[:, 0], b[:, 1] * a[:, 0], b[:, 1] * a[:, 1], b[:, 1], a[:, 0], a[:, 1],
                  np.ones(len(a))], 1)
    _, V = np.linalg.eigh((A * w[:, None]).T @ A)
    U, S, Vt = np.linalg.svd(V[:, 0].reshape(3, 3))
    F = T1.T @ (U @ np.diag([S[0], S[1], 0.0]) @ Vt) @ T0
    return F / np.linalg.norm(F)


def magsac_polish(F0: np.ndarray, m0: np.ndarray, m1: np.ndarray, cutoff: float = MAGSAC_CUTOFF_PX, max_iters: int = 200,
                  tol: float = 1e-13, hard_threshold: float = None):
    """Iterates F <- weighted_eight_point(weights(residuals(F))) to its fixed point.  hard_threshold: unit weights on
    the inliers at that threshold instead (LO-RANSAC's least squares on inliers).  Returns (F unit norm, iterations)."""
    F = F0 / np.linalg.norm(F0)
    for it in range(max_iters):
        r2 = sampson_distance(F, m0, m1) ** 2
        w = magsac_weights(r2, cutoff) if hard_threshold is None else (r2 < hard_threshold ** 2).astype(np.float64)
        Fn = weighted_eight_point(m0, m1, w)
        d = min(np.abs(Fn - F).max(), np.abs(Fn + F).max())
        F = Fn
        if d < tol:
            break
    return F, it + 1


def sampson_distance(F: np.ndarray, m0: np.ndarray, m1: np.ndarray) -> np.ndarray:
    """sqrt of the Sampson error, pixels."""
    x0 = np.concatenate([m0.astype(np.float64), np.ones((len(m0), 1))], 1)
    x1 = np.concatenate([m1.astype(np.float64), np.ones((len(m1), 1))], 1)
    Fx0 = x0 @ F.T
    Ftx1 = x1 @ F
    num = np.sum(x1 * Fx0, 1) ** 2
    den = Fx0[:, 0] ** 2 + Fx0[:, 1] ** 2 + Ftx1[:, 0] ** 2 + Ftx1[:, 1] ** 2
    return np.sqrt(num / den)


def estimate_pose(kpts0: np.ndarray, kpts1: np.ndarray, K0: np.ndarray, K1: np.ndarray, thresh: float, conf: float = 0.9999):
    """sfm/geometry.py:31-76 restated: K-normalise, cv2.findEssentialMat(RANSAC, thresh / mean focal), cv2.recoverPose with
    the in/out mask; returns (R [3,3], t [3], inliers [N] bool) of the candidate with most points in front."""
    if len(kpts0) < 5:
        return None
    f_mean = np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]])
    norm_thresh = thresh / f_mean
    kpts0 = (kpts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    kpts1 = (kpts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    E, mask = cv2.findEssentialMat(kpts0, kpts1, np.eye(3), threshold=norm_thresh, prob=conf, method=cv2.RANSAC)
    assert E is not None, "Unable to estimate Essential matrix"
    best, ret = 0, None
    for _E in np.split(E, len(E) / 3):
        n, R, t, _ = cv2.recoverPose(_E, kpts0, kpts1, np.eye(3), 1e9, mask=mask)
        if n > best:
            best, ret = n, (R, t[:, 0], mask.ravel() > 0)
    return ret


def helmert_linear(v0: np.ndarray, v1: np.ndarray, scale: bool = True) -> np.ndarray:
    """sfm/absolute_orientation.py:141-154 -> thirdparty/transformations.py:889-1020 affine_matrix_from_points(v0.T, v1.T,
    shear=False, scale=scale, usesvd=False) restated: Horn's quaternion method, scale = ratio of RMS deviations from the centroids.
    v0, v1 [n,3]; returns the 4x4 matrix with v1 ~ T v0."""
    import math
    a = np.array(v0, dtype=np.float64).T.copy()
    b = np.array(v1, dtype=np.float64).T.copy()
    t0, t1 = -a.mean(1), -b.mean(1)
    M0, M1 = np.identity(4), np.identity(4)
    M0[:3, 3], M1[:3, 3] = t0, t1
    a += t0.reshape(3, 1)
    b += t1.reshape(3, 1)
    xx, yy, zz = np.sum(a * b, axis=1)
    xy, yz, zx = np.sum(a * np.roll(b, -1, axis=0), axis=1)
    xz, yx, zy = np.sum(a * np.roll(b, -2, axis=0), axis=1)
    N = [[xx + yy + zz, 0.0, 0.0, 0.0], [yz - zy, xx - yy - zz, 0.0, 0.0], [zx - xz, xy + yx, yy - xx - zz, 0.0],
         [xy - yx, zx + xz, yz + zy, zz - xx - yy]]
    w, V = np.linalg.eigh(N)
    q = V[:, np.argmax(w)]
    q = q / np.linalg.norm(q)
    q = q * math.sqrt(2.0 / np.dot(q, q))
    o = np.outer(q, q)
    M = np.array([[1.0 - o[2, 2] - o[3, 3], o[1, 2] - o[3, 0], o[1, 3] + o[2, 0], 0.0],
                  [o[1, 2] + o[3, 0], 1.0 - o[1, 1] - o[3, 3], o[2, 3] - o[1, 0], 0.0],
                  [o[1, 3] - o[2, 0], o[2, 3] + o[1, 0], 1.0 - o[1, 1] - o[2, 2], 0.0], [0.0, 0.0, 0.0, 1.0]])
    if scale:
        M[:3, :3] *= math.sqrt(np.sum(b * b) / np.sum(a * a))
    M = np.linalg.inv(M1) @ (M @ M0)
    return M / M[3, 3]


def apply_transform(T: np.ndarray, points3d: np.ndarray) -> np.ndarray:
    """sfm/absolute_orientation.py:269-272 restated: dehomogenise(T @ [x; 1])."""
    h = T @ np.concatenate([np.asarray(points3d, dtype=np.float64).T, np.ones((1, len(points3d)))], 0)
    return (h[:3] / h[3]).T


def project_points(points3d: np.ndarray, R: np.ndarray, t: np.ndarray, K: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """sfm/geometry.py:78-100 restated: cv2.Rodrigues + cv2.projectPoints, result cast to float32."""
    rvec, _ = cv2.Rodrigues(np.asarray(R, dtype=np.float64))
    m, _ = cv2.projectPoints(np.expand_dims(np.asarray(points3d, dtype=np.float64), 1), rvec, np.asarray(t, dtype=np.float64),
                             np.asarray(K, dtype=np.float64), np.asarray(dist, dtype=np.float64))
    return m[:, 0, :].astype("float32")


def bilinear_interpolate(im: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """sfm/interpolate_colors.py:54-87 restated (neighbours are clipped before the weights are formed)."""
    x, y = np.asarray(x), np.asarray(y)
    x0 = np.floor(x).astype(int); x1 = x0 + 1
    y0 = np.floor(y).astype(int); y1 = y0 + 1
    x0 = np.clip(x0, 0, im.shape[1] - 1); x1 = np.clip(x1, 0, im.shape[1] - 1)
    y0 = np.clip(y0, 0, im.shape[0] - 1); y1 = np.clip(y1, 0, im.shape[0] - 1)
    Ia, Ib, Ic, Id = im[y0, x0], im[y1, x0], im[y0, x1], im[y1, x1]
    wa = (x1 - x) * (y1 - y); wb = (x1 - x) * (y - y0); wc = (x - x0) * (y1 - y); wd = (x - x0) * (y - y0)
    return wa * Ia + wb * Ib + wc * Ic + wd * Id


def interpolate_point_colors(points3d, image, R, t, K, dist, convert_BRG2RGB=True) -> np.ndarray:
    """sfm/interpolate_colors.py:14-51 restated."""
    assert image.ndim == 3
    if convert_BRG2RGB:
        image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
    proj = project_points(points3d, R, t, K, dist)
    image = image.astype(np.float32) / 255.0
    col = np.zeros((len(points3d), image.shape[2]))
    for ch in range(image.shape[2]):
        col[:, ch] = bilinear_interpolate(image[:, :, ch], proj[:, 0], proj[:, 1])
    return col
