"""undistort_points — drop-in for icepy4d/sfm/geometry.py:103-118 (cv2.undistortPoints(pts, K, dist, None, K) -> f32);
estimate_pose — drop-in for icepy4d/sfm/geometry.py:31-76 (cv2.findEssentialMat(RANSAC) + cv2.recoverPose)."""
import numpy as np
import torch

from .. import ops


def undistort_points(pts: np.ndarray, camera) -> np.ndarray:
    if len(pts) == 0:
        return np.zeros((0, 2), dtype="float32")
    t = torch.as_tensor(np.ascontiguousarray(pts, dtype=np.float32)).cuda()
    return ops.undistort_points(t, camera.K, camera.dist).cpu().numpy()


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.9999, max_iters=2048, seed=0):
    """Relative pose (R, t, inliers) of camera 1 w.r.t. camera 0 from matched pixel coordinates (geometry.py:31-76).

    Same normalisation and thresholds as the reference: points are K-normalised, the pixel threshold is divided by the mean
    focal length, inliers are decided by the Sampson error of the essential matrix, the (R, t) candidate is chosen by
    cheirality and the returned mask is the one recoverPose leaves behind (inliers in front of both cameras).

    Two robust estimators run on the device:
      * the five-point RANSAC the reference calls (cv2.findEssentialMat(method=RANSAC): Nister's minimal solver, models ranked by
        inlier count, OpenCV's adaptive stopping rule — csrc/essential.cu).  It is the only one for 5-7 correspondences and the
        one that survives planar / near-planar scenes, where the uncalibrated 8-point model is degenerate;
      * for >= 8 correspondences also the batched 8-point RANSAC + MAGSAC++ polish of csrc/ransac.cu on focal-scaled coordinates,
        projected onto the essential manifold (a least-squares fit on all inliers: more accurate than any minimal sample
        whenever the scene is not degenerate: it is the result unless it has >= 5 % fewer
        Sampson inliers at the caller's threshold than the best five-point model).
    OpenCV's random sample sequence is not reproduced (different minimal samples, same model class and inlier rule)."""
    if len(kpts0) < 5:
        return None
    K0, K1 = np.asarray(K0, dtype=np.float64), np.asarray(K1, dtype=np.float64)
    f_mean = float(np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]]))      # sic (geometry.py:60)
    norm_thresh = float(thresh) / f_mean
    n0 = (np.asarray(kpts0, dtype=np.float64) - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    n1 = (np.asarray(kpts1, dtype=np.float64) - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    xn0 = torch.as_tensor(np.ascontiguousarray(n0, dtype=np.float32)).cuda()
    xn1 = torch.as_tensor(np.ascontiguousarray(n1, dtype=np.float32)).cuda()
    conf = min(float(conf), 0.999999)
    cands = []
    E5, _ = ops.essential_ransac(xn0, xn1, norm_thresh, conf, int(max_iters), int(seed))
    cands.append(E5)
    if len(kpts0) >= 8:
        # x' = f_mean * x_n: the RANSAC works in pixel-like units with the caller's pixel threshold; E = D F' D, D = diag(f, f, 1)
        Fs, _, _ = ops.fundamental_ransac((xn0 * f_mean).contiguous(), (xn1 * f_mean).contiguous(), float(thresh), conf, 10000, seed)
        D = torch.tensor([f_mean, f_mean, 1.0], dtype=torch.float64, device="cuda")
        cands.append((D[:, None] * Fs.view(3, 3) * D[None, :]).reshape(9))
    best = None
    for E in cands:
        if not bool(torch.isfinite(E).all()) or float(E.abs().max()) == 0.0:
            continue
        _, R, t, mask, n_good = ops.essential_pose(E, xn0, xn1, norm_thresh, 1e9)
        n_inl = int(n_good[1])
        # the polished model (second) is a least-squares fit on all its inliers and replaces the best minimal sample unless it
        # has clearly lost support (>= 5 % fewer inliers: the signature of a degenerate 8-point configuration)
        if best is None or n_inl >= 0.95 * best[0]:
            best = (n_inl, R, t, mask)
    assert best is not None, "Unable to estimate Essential matrix"          # geometry.py:67
    _, R, t, mask = best
    return R.view(3, 3).cpu().numpy(), t.cpu().numpy(), mask.cpu().numpy() > 0
