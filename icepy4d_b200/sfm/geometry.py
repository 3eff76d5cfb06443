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


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.9999):
    """Relative pose (R, t, inliers) of camera 1 w.r.t. camera 0 from matched pixel coordinates (geometry.py:31-76).

    Same normalisation and thresholds as the reference: points are K-normalised, the pixel threshold is divided by the mean
    focal length, inliers are decided by the Sampson error of the essential matrix, the (R, t) candidate is chosen by
    cheirality and the returned mask is the one recoverPose leaves behind (inliers in front of both cameras).
    The robust estimate comes from the batched-hypothesis CUDA RANSAC (ransac.cu) run on focal-scaled normalised coordinates
    — its 8-point hypotheses live in the calibrated space, so the model it returns is an essential matrix up to the manifold
    projection done on the device (pose.cu); OpenCV's 5-point minimal solver is not reproduced (different sampler, same
    model class and inlier rule)."""
    if len(kpts0) < 5:
        return None
    K0, K1 = np.asarray(K0, dtype=np.float64), np.asarray(K1, dtype=np.float64)
    f_mean = float(np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]]))      # sic (geometry.py:60)
    norm_thresh = float(thresh) / f_mean
    n0 = (np.asarray(kpts0, dtype=np.float64) - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    n1 = (np.asarray(kpts1, dtype=np.float64) - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    xn0 = torch.as_tensor(np.ascontiguousarray(n0, dtype=np.float32)).cuda()
    xn1 = torch.as_tensor(np.ascontiguousarray(n1, dtype=np.float32)).cuda()
    if len(kpts0) >= 8:
        # x' = f_mean * x_n: the RANSAC works in pixel-like units with the caller's pixel threshold; E = D F' D, D = diag(f, f, 1)
        Fs, _, _ = ops.fundamental_ransac((xn0 * f_mean).contiguous(), (xn1 * f_mean).contiguous(), float(thresh),
                                          min(float(conf), 0.999999), 10000, 0)
        D = torch.tensor([f_mean, f_mean, 1.0], dtype=torch.float64, device="cuda")
        E = (D[:, None] * Fs.view(3, 3) * D[None, :]).reshape(9)
    else:
        raise ValueError("estimate_pose needs at least 8 correspondences on the B200 path")
    assert bool(torch.isfinite(E).all()) and float(E.abs().max()) > 0, "Unable to estimate Essential matrix"   # geometry.py:74
    _, R, t, mask, n_good = ops.essential_pose(E, xn0, xn1, norm_thresh, 1e9)
    return R.view(3, 3).cpu().numpy(), t.cpu().numpy(), mask.cpu().numpy() > 0
