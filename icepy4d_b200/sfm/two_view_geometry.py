"""RelativeOrientation.estimate_F_matrix — drop-in for icepy4d/sfm/two_view_geometry.py:127-197.

Returns (F [3,3] f64, inlMask [N] bool) and replaces `self.features[i]` by the inliers, like the reference.  The
reference's MAGSAC branch raises IndexError because its mask is [N,1] (Appendix D.6); here the mask is [N] bool as in
`geometric_verification`.  `estimate_pose` follows two_view_geometry.py:52-109 on top of sfm.geometry.estimate_pose."""
from __future__ import annotations

import logging
from typing import List

import numpy as np

from ..matching.enums import GeometricVerification
from ..matching.geometric_verification import geometric_verification


class RelativeOrientation:
    def __init__(self, cameras: List, features: List[np.ndarray]) -> None:
        self.cameras = cameras
        self.features = features
        self.F = None
        self.inlMask = None

    def estimate_F_matrix(self, threshold: float = 1, confidence: float = 0.9999, max_iters: int = 10000,
                          laf_consistensy_coef: float = -1.0, error_type: str = "sampson",
                          symmetric_error_check: bool = True, enable_degeneracy_check: bool = True):
        # two_view_geometry.py:165-175: the pydegensac branch honours the caller's threshold / confidence / max_iters (the
        # MAGSAC fallback of :180-187, taken when pydegensac is missing, replaces them by 0.5 px / 0.999 / 100000 and then
        # raises IndexError, Appendix D.6).  The arguments are part of this method's contract, so they take effect here.
        self.F, self.inlMask = geometric_verification(self.features[0], self.features[1], GeometricVerification.PYDEGENSAC,
                                                      threshold, confidence, max_iters)
        logging.info(f"found {self.inlMask.sum()} inliers ({self.inlMask.sum() * 100 / max(1, len(self.features[0])):.2f}%)")
        self.features[0] = self.features[0][self.inlMask]
        self.features[1] = self.features[1][self.inlMask]
        return (self.F, self.inlMask)

    def estimate_pose(self, threshold: float = 1.0, confidence: float = 0.9999, scale_factor=None):
        """two_view_geometry.py:52-109: essential-matrix relative orientation, then camera 1's extrinsics are replaced by the
        estimated ones chained with camera 0's pose (same Camera calls as the reference, so its Camera class works here)."""
        from .geometry import estimate_pose

        assert self.cameras[0].extrinsics is not None, \
            "Extrinsics matrix is not available for camera 0. Please, compute it before running RelativeOrientation estimation."
        R, t, valid = estimate_pose(self.features[0], self.features[1], self.cameras[0].K, self.cameras[1].K,
                                    thresh=threshold, conf=confidence)
        logging.info(f"Relative Orientation - valid points: {valid.sum()}/{len(valid)}")
        if scale_factor is not None:
            t = t * scale_factor
        else:
            logging.warning("No scaling factor (e.g., computed from camera baseline) is provided. "
                            "Two-view-geometry estimated up to a scale factor.")
        extrinsics = self.cameras[1].Rt_to_extrinsics(R, t)
        self.cameras[1].update_extrinsics(extrinsics)
        cam2toWorld = self.cameras[0].pose @ self.cameras[1].pose
        extrinsics = self.cameras[1].pose_to_extrinsics(cam2toWorld)
        self.cameras[1].update_extrinsics(extrinsics)
        logging.info("Relative orientation Succeded.")
        return valid

    def get_scale_factor_from_baseline(self, baseline_world: float):
        """two_view_geometry.py:111-125"""
        baseline = np.linalg.norm(np.asarray(self.cameras[0].C) - np.asarray(self.cameras[1].C))
        return baseline_world / baseline
