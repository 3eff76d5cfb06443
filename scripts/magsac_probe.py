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

  (4) --translation: the matcher golden (tests/golden/matchers.npz) is a pure image translation; a 2-parameter family of
      rank-2 F fits all its true matches exactly, so the MAGSAC++ objective has several local optima that differ only in
      which WRONG matches they accept.  RANSAC (8-point samples, best of 300 by MAGSAC++ quality) + polisher from 30 seeds:
      the optimum most seeds reach has a HIGHER quality than the model cv2 returns and an inlier IoU of 0.95 with it.

    python scripts/magsac_probe.py [n ...]            # default 20000 50000
    python scripts/magsac_probe.py --translation
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


def translation_landscape():
    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "matchers.npz"))
    p0, p1 = d["sg_mkpts0"], d["sg_mkpts1"]
    n = len(p0)
    true = (p0[:, 0] - p1[:, 0] == 16) & (p0[:, 1] - p1[:, 1] == 8)
    runs = cv2_runs(p0, p1, 5)
    print(f"matcher golden: {n} matches, {true.sum()} true; cv2 runs keep {[int(m.sum()) for _, m in runs]} "
          f"(all true matches: {[bool((m & true).sum() == true.sum()) for _, m in runs]}), MAGSAC++ quality of cv2's models "
          f"{[round(g.magsac_quality(F, p0, p1), 2) for F, _ in runs]}")
    rng = np.random.default_rng(0)
    rows = {}
    for trial in range(30):
        best, bq = None, -1.0
        for _ in range(300):
            idx = rng.choice(n, 8, replace=False)
            F = g.weighted_eight_point(p0[idx], p1[idx], np.ones(8))
            if not np.isfinite(F).all():
                continue
            q = g.magsac_quality(F, p0, p1)
            if q > bq:
                best, bq = F, q
        Fp, _ = g.magsac_polish(best, p0, p1)
        mp = g.sampson_distance(Fp, p0, p1) < 0.5
        key = (round(g.magsac_quality(Fp, p0, p1), 2), int(mp.sum()), round(iou(mp, runs[0][1]), 4), bool((mp & true).sum() == true.sum()))
        rows[key] = rows.get(key, 0) + 1
    print("   (quality after polish, inliers, IoU vs cv2, keeps all true matches) -> number of seeds")
    for k in sorted(rows, reverse=True):
        print("  ", k, "->", rows[k])


def main():
    if "--translation" in sys.argv:
        print(f"OpenCV {cv2.__version__}")
        return translation_landscape()
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
