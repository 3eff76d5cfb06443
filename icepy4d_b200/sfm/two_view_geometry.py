"""RelativeOrientation.estimate_F_matrix — drop-in for icepy4d/sfm/two_view_geometry.py:127-197.

Returns (F [3,3] f64, inlMask [N] bool) and replaces `self.features[i]` by the inliers, like the reference.  The
reference's MAGSAC branch raises IndexError because its mask is [N,1] (Appendix D.6); here the mask is [N] bool as in
`geometric_verification`.  `estimate_pose` (5-point essential matrix) is the "next" row of SURVEY.md §8f."""
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
        # pydegensac is absent here as it is in the reference's own fallback: the reference then runs MAGSAC with
        # its hard-coded parameters (two_view_geometry.py:180-187)
        self.F, self.inlMask = geometric_verification(self.features[0], self.features[1], GeometricVerification.MAGSAC,
                                                      threshold, confidence, max_iters)
        logging.info(f"found {self.inlMask.sum()} inliers ({self.inlMask.sum() * 100 / max(1, len(self.features[0])):.2f}%)")
        self.features[0] = self.features[0][self.inlMask]
        self.features[1] = self.features[1][self.inlMask]
        return (self.F, self.inlMask)

    def estimate_pose(self, *a, **k):
        raise NotImplementedError("estimate_pose (essential-matrix RANSAC + recoverPose) is outside round 1 (SURVEY.md §8f rank 1)")
