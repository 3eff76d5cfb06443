"""undistort_points — drop-in for icepy4d/sfm/geometry.py:103-118 (cv2.undistortPoints(pts, K, dist, None, K) -> f32)."""
import numpy as np
import torch

from .. import ops


def undistort_points(pts: np.ndarray, camera) -> np.ndarray:
    if len(pts) == 0:
        return np.zeros((0, 2), dtype="float32")
    t = torch.as_tensor(np.ascontiguousarray(pts, dtype=np.float32)).cuda()
    return ops.undistort_points(t, camera.K, camera.dist).cpu().numpy()
