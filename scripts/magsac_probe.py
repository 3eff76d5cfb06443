"""What does cv2.findFundamentalMat(USAC_MAGSAC, 0.5, 0.999, 100000) — the reference's runnable verification branch
(matching/geometric_verification.py:89-92) — pin down, and where does the B200 polisher have to land?  CPU only.

  (1) self-consistency: the same call on the SAME correspondences in a different order (OpenCV's sampler is seeded with a
      constant, so a permutation is the only way to see its sampling noise).  On cfg3 (Brown distortion left in the pixels,
      0.5 px threshold) its inlier sets agree with each other only to IoU 0.95-0.99: the returned F differ by ~1e-6 and
      the mask is a knife-edge function of F.  No estimator can agree with every one of those runs better than they agree
      with each other.
  (2) cut-off fit: OpenCV's final model is (nearly) a fixed point of the MAGSAC++ re-weighted 8-point fit; scanning the
      cut-off k*sigma_max shows the minimum one-step displacement at ~4.5 px, independent of the user threshold.
  (3) the fixed point of that polisher (oracle/geom_oracle.py::magsac_polish, which csrc/ransac.cu implements) against the
      cv2 runs of (1): it agrees with them as well as they agree with each other.

    python scripts/magsac_probe.py [n ...]            # default 20000 50000
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402
import numpy as np  # noqa: E402

from icepy4d_b200 import synthetic  # noqa: E402
from oracle import geom_oracle as g  # noqa: E402


def iou(a, b):
    return (a & b).sum() / max(1, (a | b).sum())


def cv2_runs(p0, p1, k, seed=1):
    rng = np.random.default_rng(seed)
    n = len(p0)
    out = []
    for t in range(k):
        perm = rng.permutation(n) if t else np.arange(n)
        F, m = cv2.findFundamentalMat(p0[perm], p1[perm], cv2.USAC_MAGSAC, 0.5, 0.999, 100000)
        mu = np.zeros(n, bool)
        mu[perm] = m.ravel() > 0
        out.append((F / np.linalg.norm(F), mu))
    return out


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [20000, 50000]
    print(f"OpenCV {cv2.__version__}")
    print("(2) one-step displacement of cv2's F under the MAGSAC++ polisher, by cut-off [px] (distortion-free, no outliers)")
    for noise, thr in ((1.0, 0.5), (2.0, 0.5), (1.0, 1.0), (1.0, 2.0)):
        sc = synthetic.two_view_scene(n=20000, seed=11, noise_px=noise, outlier_frac=0.0, distortion=False)
        F, _ = cv2.findFundamentalMat(sc["pts0"], sc["pts1"], cv2.USAC_MAGSAC, thr, 0.999, 100000)
        F = F / np.linalg.norm(F)
        r2 = g.sampson_distance(F, sc["pts0"], sc["pts1"]) ** 2
        row = []
        for c in (2.5, 3.5, 4.0, 4.25, 4.5, 4.75, 5.0, 6.0, 8.0):
            Fn = g.weighted_eight_point(sc["pts0"], sc["pts1"], g.magsac_weights(r2, c))
            row.append(f"{c}:{min(np.abs(Fn - F).max(), np.abs(Fn + F).max()):.1e}")
        print(f"   noise {noise} px, threshold {thr} px   " + "  ".join(row))
    for n in sizes:
        sc = synthetic.two_view_scene(n=n, seed=7)
        runs = cv2_runs(sc["pts0"], sc["pts1"], 5)
        self_iou = [iou(a[1], b[1]) for a, b in itertools.combinations(runs, 2)]
        print(f"(1) cfg3 n={n}: cv2 vs cv2 (5 orders of the same points) inlier IoU mean {np.mean(self_iou):.4f} "
              f"min {np.min(self_iou):.4f} max {np.max(self_iou):.4f}; "
              f"max |dF| between runs {max(min(np.abs(a[0] - b[0]).max(), np.abs(a[0] + b[0]).max()) for a, b in itertools.combinations(runs, 2)):.1e}")
        for c in (3.64, 4.0, 4.5, 5.0):
            F, it = g.magsac_polish(runs[0][0], sc["pts0"], sc["pts1"], c)
            m = g.sampson_distance(F, sc["pts0"], sc["pts1"]) < 0.5
            ious = [iou(m, r[1]) for r in runs]
            print(f"(3)    polisher fixed point, cut-off {c} px ({it} iterations): IoU vs the cv2 runs mean {np.mean(ious):.4f} "
                  f"min {np.min(ious):.4f}")


if __name__ == "__main__":
    main()
