"""BASELINE.json configs as GPU parity / property tests at their full sizes (configs[1] is the bench line):
cfg1 (LightGlue, 2048 kp, 1500x1000) against the CPU oracle, cfg3 (200 k correspondences) against OpenCV + round-trip
properties, cfg5 (16384 kp LightGlue tile pair) through size-independent properties of the assignment."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from icepy4d_b200 import synthetic, weights  # noqa: E402
from oracle import geom_oracle, lg_oracle, sp_oracle  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


def test_cfg1_lightglue_2048kp_1500x1000_vs_oracle():
    from icepy4d_b200.matching import GeometricVerification, LightGlueMatcher, Quality, TileSelection
    i0, i1 = synthetic.stereo_pair(1000, 1500, seed=1003, shift=(24, 16), channels=3)
    sp_sd, lg_sd = weights.make_superpoint_state(1), weights.make_lightglue_state(3)
    m = LightGlueMatcher({"features": "superpoint", "superpoint_state": sp_sd, "lightglue_state": lg_sd, "precision": "bf16",
                          "conv_precision": "f32"})
    m.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[1, 1], overlap=0, max_keypoints=2048,
            geometric_verification=GeometricVerification.NONE)
    got = {(float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(m.mkpts0, m.mkpts1)}
    # oracle: the [1,1] "tile" is 1499 x 999 (end-exclusive slicing), grey = 0.299 r + 0.587 g + 0.114 b in f32
    def grey(img):
        t = torch.tensor(img[:999, :1499].transpose(2, 0, 1) / 255.0, dtype=torch.float)
        return (t * torch.tensor([0.299, 0.587, 0.114]).view(3, 1, 1)).sum(0)[None, None]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    f0, f1 = sp_oracle.superpoint_lg(grey(i0), sp_sd, k=2048), sp_oracle.superpoint_lg(grey(i1), sp_sd, k=2048)
    out = lg_oracle.lightglue(f0["keypoints"], f0["descriptors"], (1499.0, 999.0), f1["keypoints"], f1["descriptors"], (1499.0, 999.0), lg_sd)
    mt = out["matches"].numpy()
    k0, k1 = f0["keypoints"].numpy()[mt[:, 0]], f1["keypoints"].numpy()[mt[:, 1]]
    ref = {(float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(k0, k1)}
    iou = len(got & ref) / max(1, len(got | ref))
    print("cfg1 LightGlue match IoU vs oracle:", iou, len(got), len(ref))
    assert iou >= 0.99, iou
    shift_ok = np.mean((m.mkpts0[:, 0] - m.mkpts1[:, 0] == 24) & (m.mkpts0[:, 1] - m.mkpts1[:, 1] == 16))
    assert shift_ok > 0.85                                         # ground truth: image1(x, y) = image0(x + 24, y + 16)


def test_cfg3_200k_verification_and_triangulation():
    from icepy4d_b200 import ops
    from icepy4d_b200.sfm import Triangulate
    sc = synthetic.two_view_scene(n=200_000, seed=11)
    p0, p1 = torch.from_numpy(sc["pts0"]).cuda(), torch.from_numpy(sc["pts1"]).cuda()
    F, mask, n = ops.fundamental_ransac(p0, p1, 0.5, 0.999, 100000, seed=0)
    mask = mask.cpu().numpy().astype(bool)
    Fh = F.cpu().numpy().reshape(3, 3)
    d = geom_oracle.sampson_distance(Fh, sc["pts0"], sc["pts1"])
    assert ((d < 0.5) == mask).mean() > 0.99999                    # inlier rule = OpenCV USAC's (sqrt Sampson < thr)
    assert abs(np.linalg.det(Fh / np.linalg.norm(Fh))) < 1e-10     # rank 2
    assert (mask & ~sc["inlier"]).sum() < 0.002 * mask.sum()       # gross outliers rejected
    assert mask[sc["inlier"]].mean() > 0.45                        # distorted pixels: ~half of the true inliers are within 0.5 px
    # cv2 USAC_MAGSAC (the reference's runnable branch) on the same points in three orders: it agrees with itself to
    # 0.98-0.999 on this config (knife-edge inlier rule under a misfitting pinhole model, tests/test_gpu_plugin.py), so the
    # gates are (a) >= 0.999 against the CPU restatement of the MAGSAC++ polisher started from cv2's model and (b) at least
    # as close to the cv2 runs as they are to each other.
    import cv2
    rng = np.random.default_rng(1)
    cvm, Fr = [], None
    for t in range(3):
        perm = rng.permutation(len(mask)) if t else np.arange(len(mask))
        Fc, m = cv2.findFundamentalMat(sc["pts0"][perm], sc["pts1"][perm], cv2.USAC_MAGSAC, 0.5, 0.999, 100000)
        Fr = Fc if Fr is None else Fr
        mu = np.zeros(len(mask), bool)
        mu[perm] = m.ravel() > 0
        cvm.append(mu)
    iou = lambda a, b: (a & b).sum() / max(1, (a | b).sum())
    self_iou = [iou(cvm[0], cvm[1]), iou(cvm[0], cvm[2]), iou(cvm[1], cvm[2])]
    ours = [iou(mask, m) for m in cvm]
    Fo, _ = geom_oracle.magsac_polish(Fr, sc["pts0"], sc["pts1"])
    mo = geom_oracle.sampson_distance(Fo, sc["pts0"], sc["pts1"]) < 0.5
    Fn = Fh / np.linalg.norm(Fh)
    dF = min(np.abs(Fn - Fo).max(), np.abs(Fn + Fo).max())
    print(f"cfg3 200k: inlier IoU vs oracle polisher {iou(mask, mo):.4f} (|dF| {dF:.1e}); vs cv2 runs {np.round(ours, 4)}; "
          f"cv2 vs cv2 {np.round(self_iou, 4)}")
    assert iou(mask, mo) >= 0.999 and dF < 1e-9
    assert np.mean(ours) >= np.mean(self_iou) - 0.01 and min(ours) >= min(self_iou) - 0.01
    # triangulation of the true inliers: round trip through the known cameras
    inl = sc["inlier"]
    tri = Triangulate(sc["cams"], [sc["pts0"][inl], sc["pts1"][inl]])
    X = tri.triangulate_two_views()
    assert X.shape == (inl.sum(), 3) and (tri.status == 1).mean() > 0.999
    rel = np.linalg.norm(X - sc["X"][inl], axis=1) / np.linalg.norm(sc["X"][inl], axis=1)
    assert np.median(rel) < 5e-4                                   # 0.3 px noise at 600 m
    # against the oracle on a sample (north-star tolerance 1e-4, measured ~1e-12)
    idx = np.random.default_rng(0).choice(inl.sum(), 300, replace=False)
    u0 = geom_oracle.undistort_points(sc["pts0"][inl][idx], sc["cams"][0].K, sc["cams"][0].dist)
    u1 = geom_oracle.undistort_points(sc["pts1"][inl][idx], sc["cams"][1].K, sc["cams"][1].dist)
    Xo, so = geom_oracle.iterative_ls(u0, sc["cams"][0].P, u1, sc["cams"][1].P)
    assert (np.linalg.norm(X[idx] - Xo, axis=1) / np.linalg.norm(Xo, axis=1)).max() < 1e-8
    Xl = Triangulate(sc["cams"], [sc["pts0"][inl][idx], sc["pts1"][inl][idx]]).triangulate_two_views(approach="linear_triangulation")
    Xlo = geom_oracle.linear_dlt(u0, sc["cams"][0].P, u1, sc["cams"][1].P)
    assert (np.linalg.norm(Xl - Xlo, axis=1) / np.linalg.norm(Xlo, axis=1)).max() < 1e-7


def test_cfg5_lightglue_16384kp_tile_pair_properties():
    """One 1499x1329 tile pair at 16384 kp/tile (static depth/width): the oracle would need minutes and ~10 GB here, so the
    full size is checked through properties of the assignment; parity at oracle-sized inputs is in test_gpu_plugin/tc."""
    from icepy4d_b200.matching.lightglue import LightGlueB200
    from icepy4d_b200.matching.superpoint import SuperPointB200, sync_counts
    i0, i1 = synthetic.stereo_pair(1329, 1499, seed=1005, shift=(16, 8), channels=1)
    sp = SuperPointB200(weights.make_superpoint_state(1), nms_radius=4, keypoint_threshold=0.0005, max_keypoints=16384, conv_precision="tf32")
    lg = LightGlueB200(weights.make_lightglue_state(3), precision="bf16", depth_confidence=-1, width_confidence=-1)
    t = lambda a: torch.tensor(a / 255.0, dtype=torch.float)[None, None].cuda()
    f0, f1 = sp.detect(t(i0)), sp.detect(t(i1))
    sync_counts(f0, f1)
    assert f0.n == 16384 and f1.n == 16384
    out = lg.match(f0.keypoints, f0.descriptors, (1499, 1329), f1.keypoints, f1.descriptors, (1499, 1329))
    m0, m1 = out["matches0"].cpu().numpy(), out["matches1"].cpu().numpy()
    s0 = out["matching_scores0"].cpu().numpy()
    v = m0 > -1
    assert v.sum() > 4000
    assert np.array_equal(m1[m0[v]], np.nonzero(v)[0])             # mutual consistency
    assert len(set(m0[v].tolist())) == v.sum()                     # one-to-one
    assert np.all(s0[v] > 0.1) and np.all(s0 <= 1.0 + 1e-5) and np.all(s0[~v] <= 0.1 + 1e-6) 
    k0, k1 = f0.keypoints.cpu().numpy()[v], f1.keypoints.cpu().numpy()[m0[v]]
    ok = np.mean((k0[:, 0] - k1[:, 0] == 16) & (k0[:, 1] - k1[:, 1] == 8))
    print("cfg5 16384-kp tile pair: matches", v.sum(), "consistent with the ground-truth shift:", ok)
    assert ok > 0.9
