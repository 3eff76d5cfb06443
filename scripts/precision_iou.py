"""End-to-end match-set IoU of the plugin vs the reference's golden output for every precision policy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icepy4d_b200 import weights
from icepy4d_b200.matching import SuperGlueMatcher, LightGlueMatcher, GeometricVerification, Quality, TileSelection
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/matchers.npz"))
def pairs(a, b): return {(float(p[0]), float(p[1]), float(q[0]), float(q[1])) for p, q in zip(a, b)}
def kset(a): return {(float(p[0]), float(p[1])) for p in a}
ref = pairs(g["sg_mkpts0"], g["sg_mkpts1"]); refk = kset(g["sg_mkpts0"])
refl = pairs(g["lg_mkpts0"], g["lg_mkpts1"])
for conv in (tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else ("bf16x3", "f16x3", "f32", "tf32", "f16", "bf16")):
    for prec in ("f32", "bf16"):
        m = SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": 512, "match_threshold": 0.2,
                              "force_cpu": False, "sinkhorn_iterations": 20, "superpoint_state": weights.make_superpoint_state(1),
                              "superglue_state": weights.make_superglue_state(2), "precision": prec, "conv_precision": conv})
        m.match(g["image0"], g["image1"], quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
                geometric_verification=GeometricVerification.NONE)
        a = pairs(m.mkpts0, m.mkpts1); ak = kset(m.mkpts0)
        l = LightGlueMatcher({"features": "superpoint", "superpoint_state": weights.make_superpoint_state(1),
                              "lightglue_state": weights.make_lightglue_state(3), "precision": prec, "conv_precision": conv})
        l.match(g["image0"], g["image1"], quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
                max_keypoints=512, geometric_verification=GeometricVerification.NONE)
        b = pairs(l.mkpts0, l.mkpts1)
        print(f"conv={conv:5s} matcher={prec:5s}  SG match IoU {len(a&ref)/len(a|ref):.4f} (n={len(a)})  matched-kpt IoU {len(ak&refk)/len(ak|refk):.4f}   LG match IoU {len(b&refl)/len(b|refl):.4f} (n={len(b)})", flush=True)
