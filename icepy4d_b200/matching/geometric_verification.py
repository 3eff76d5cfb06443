"""geometric_verification — same signature / return convention as the reference
(icepy4d/matching/geometric_verification.py:11-102), computed by the batched-hypothesis CUDA RANSAC.

Reference semantics kept: fewer than 4 matches -> (None, all-True); the MAGSAC branch ignores the caller's
threshold/confidence/max_iters and uses (0.5 px, 0.999, 100000) (geometric_verification.py:89-91, Appendix D.5) with the
MAGSAC++ quality function and polisher of cv2.USAC_MAGSAC; exceptions degrade to an all-inlier mask with F = None and are
logged, never raised (:96-100); degenerate input (no valid model) gives (None, all-False) as cv2 does; 4-7 matches give
(None, all-True) (cv2 raises below 7 points; its 7-point model for exactly 7 matches is not reproduced).

PYDEGENSAC branch (the reference's default, :66-76): pydegensac is not installed in the reference's own environment here and
its source is not vendored, so its output cannot be pinned.  What is matched is its documented behaviour: LO-RANSAC on the
Sampson error with the CALLER's px threshold, confidence and max_iters, the model being the least-squares fit on its own
inliers (`polish_mode=1`).  Not reproduced: DEGENSAC's plane-degeneracy (H-consistency) test, the LAF check and the symmetric
error check — the arguments are accepted and ignored.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import ops
from .enums import GeometricVerification

logger = logging.getLogger(__name__)

MAGSAC_PARAMS = (0.5, 0.999, 100000)


def geometric_verification_device(mkpts0: torch.Tensor, mkpts1: torch.Tensor,
                                  method: GeometricVerification = GeometricVerification.PYDEGENSAC,
                                  threshold: float = 1, confidence: float = 0.9999, max_iters: int = 10000, seed: int = 0):
    """Device-resident variant: returns (F [9] f64 device or None, mask [n] bool device).  No host sync."""
    n = mkpts0.shape[0]
    if n < 8:
        return None, torch.ones(n, dtype=torch.bool, device=mkpts0.device)
    polish_mode = 1
    if method == GeometricVerification.MAGSAC:
        threshold, confidence, max_iters = MAGSAC_PARAMS
        polish_mode = 0
    F, mask, _ = ops.fundamental_ransac(mkpts0.contiguous(), mkpts1.contiguous(), threshold, min(confidence, 0.999999),
                                        max_iters, seed, polish_mode=polish_mode)
    return F, mask.bool()


def geometric_verification(mkpts0: np.ndarray = None, mkpts1: np.ndarray = None,
                           method: GeometricVerification = GeometricVerification.PYDEGENSAC, threshold: float = 1,
                           confidence: float = 0.9999, max_iters: int = 10000, laf_consistensy_coef: float = -1.0,
                           error_type: str = "sampson", symmetric_error_check: bool = True,
                           enable_degeneracy_check: bool = True):
    assert isinstance(method, GeometricVerification), "Invalid method. It must be a GeometricVerification enum."
    F = None
    inl = np.ones(len(mkpts0), dtype=bool)
    if len(mkpts0) < 4:
        logger.warning("Not enough matches to perform geometric verification.")
        return F, inl
    try:
        a = torch.as_tensor(np.ascontiguousarray(mkpts0, dtype=np.float32)).cuda()
        b = torch.as_tensor(np.ascontiguousarray(mkpts1, dtype=np.float32)).cuda()
        Fd, mask = geometric_verification_device(a, b, method, threshold, confidence, max_iters)
        if Fd is not None:
            F = Fd.cpu().numpy().reshape(3, 3)
            inl = mask.cpu().numpy()
            if not np.isfinite(F).all():        # degenerate input: cv2 returns (None, zeros) and the reference passes that on (:89-92)
                logger.error("RANSAC found no model.")
                return None, inl
            logger.info(f"B200 RANSAC found {inl.sum()} inliers ({inl.sum() * 100 / len(mkpts0):.2f}%)")
    except Exception as err:  # same degrade-don't-raise convention as the reference (:96-100)
        logger.error(f"{err}. Unable to perform geometric verification.")
        inl = np.ones(len(mkpts0), dtype=bool)
    return F, inl
