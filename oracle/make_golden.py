"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference's own tests hold no golden vectors for this path, so these fixtures — outputs
of the reference code itself on seeded inputs with the seeded weights of icepy4d_b200/weights.py — are what
pins the oracle (tests/test_oracle_vs_golden.py) and, through it, the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from icepy4d_b200 import synthetic, weights  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _t(img):
    return torch.tensor(img / 255.0, dtype=torch.float)[None, None]


def golden_superpoint_superglue():
    sp_sd, sg_sd = weights.make_superpoint_state(1), weights.make_superglue_state(2)
    i0, i1 = synthetic.stereo_pair(240, 320, seed=1000, shift=(16, 24), channels=1)
    sp = ref_shims.build_reference_superpoint_sg({"nms_radius": 3, "keypoint_threshold": 1e-4, "max_keypoints": 256}, sp_sd)
    sg = ref_shims.build_reference_superglue({"weights": "outdoor", "sinkhorn_iterations": 20, "match_threshold": 0.2}, sg_sd)
    from icepy4d.thirdparty.SuperGlue.models.superpoint import simple_nms
    with torch.inference_mode():
        p0, p1 = sp({"image": _t(i0)}), sp({"image": _t(i1)})
        # the NMS'd score map itself (reference function applied to the reference's own score tensor)
        x = _t(i0)
        for a, b in (("conv1a", "conv1b"), ("conv2a", "conv2b"), ("conv3a", "conv3b"), ("conv4a", "conv4b")):
            x = sp.relu(getattr(sp, b)(sp.relu(getattr(sp, a)(x))))
            if a != "conv4a":
                x = sp.pool(x)
        logits = sp.convPb(sp.relu(sp.convPa(x)))
        sc = torch.softmax(logits, 1)[:, :-1]
        b_, _, h, w = sc.shape
        sc = sc.permute(0, 2, 3, 1).reshape(b_, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(b_, h * 8, w * 8)
        nms = simple_nms(sc, 3)[0]
        data = {"image0": _t(i0), "image1": _t(i1),
                "keypoints0": torch.stack(p0["keypoints"]), "keypoints1": torch.stack(p1["keypoints"]),
                "scores0": torch.stack(p0["scores"]), "scores1": torch.stack(p1["scores"]),
                "descriptors0": torch.stack(p0["descriptors"]), "descriptors1": torch.stack(p1["descriptors"])}
        out = sg(data)
    np.savez_compressed(
        os.path.join(OUT, "sp_sg_small.npz"), image0=i0, image1=i1, logits0=logits[0].numpy(),
        nms0=nms.numpy(),
        kpts0=p0["keypoints"][0].numpy(), kpts1=p1["keypoints"][0].numpy(),
        scores0=p0["scores"][0].numpy(), scores1=p1["scores"][0].numpy(),
        desc0=p0["descriptors"][0].numpy(), desc1=p1["descriptors"][0].numpy(),
        matches0=out["matches0"][0].numpy(), matches1=out["matches1"][0].numpy(),
        mscores0=out["matching_scores0"][0].numpy(), mscores1=out["matching_scores1"][0].numpy())
    print("sp_sg_small: kpts", len(p0["keypoints"][0]), len(p1["keypoints"][0]), "matches", int((out["matches0"] > -1).sum()))


def golden_lightglue():
    sp_sd = weights.make_superpoint_state(1)
    i0, i1 = synthetic.stereo_pair(240, 320, seed=1001, shift=(8, 16), channels=1)
    for tag, conf_layers, prune, kw in (("plain", (), False, {}), ("adaptive", (2, 3, 4, 5, 6, 7), False, {}),
                                        ("prune", (), True, {"depth_confidence": -1})):
        lg_sd = weights.make_lightglue_state(3, confident_layers=conf_layers, prune_variant=prune)
        sp = ref_shims.build_reference_superpoint_lg(sp_sd, max_num_keypoints=256)
        lg = ref_shims.build_reference_lightglue(lg_sd, **kw)
        with torch.inference_mode():
            f0 = sp.extract(torch.tensor(i0 / 255.0, dtype=torch.float)[None], resize=None)
            f1 = sp.extract(torch.tensor(i1 / 255.0, dtype=torch.float)[None], resize=None)
            out = lg({"image0": f0, "image1": f1})
        np.savez_compressed(
            os.path.join(OUT, f"lg_{tag}.npz"), image0=i0, image1=i1,
            kpts0=f0["keypoints"][0].numpy(), kpts1=f1["keypoints"][0].numpy(),
            scores0=f0["keypoint_scores"][0].numpy(), scores1=f1["keypoint_scores"][0].numpy(),
            desc0=f0["descriptors"][0].numpy(), desc1=f1["descriptors"][0].numpy(),
            size0=f0["image_size"][0].numpy(), size1=f1["image_size"][0].numpy(),
            matches0=out["matches0"][0].numpy(), matches1=out["matches1"][0].numpy(),
            mscores0=out["matching_scores0"][0].numpy(), mscores1=out["matching_scores1"][0].numpy(),
            matches=out["matches"][0].numpy(), scores=out["scores"][0].numpy(), stop=np.int64(out["stop"]),
            prune0=out["prune0"][0].numpy(), prune1=out["prune1"][0].numpy(),
            confident_layers=np.array(conf_layers, dtype=np.int64), prune_variant=np.bool_(prune),
            depth_confidence=np.float64(kw.get("depth_confidence", 0.95)))
        print(f"lg_{tag}: kpts", f0["keypoints"].shape[1], f1["keypoints"].shape[1], "matches",
              len(out["matches"][0]), "stop", out["stop"],
              "prune0 min/max", int(out["prune0"].min()), int(out["prune0"].max()))


def golden_matchers():
    """Whole-plugin goldens: SuperGlueMatcher.match / LightGlueMatcher.match with GRID tiling + MAGSAC."""
    ref_shims.install_shims()
    import icepy4d.matching.matchers as M
    from icepy4d.matching import GeometricVerification, Quality, TileSelection

    sp_sd, sg_sd, lg_sd = weights.make_superpoint_state(1), weights.make_superglue_state(2), weights.make_lightglue_state(3)
    i0, i1 = synthetic.stereo_pair(480, 640, seed=1002, shift=(16, 8), channels=3)
    with ref_shims.no_checkpoint_loading():
        m = M.SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": 512,
                                "match_threshold": 0.2, "force_cpu": True, "sinkhorn_iterations": 20})
    m.matcher.superpoint.load_state_dict(sp_sd)
    m.matcher.superglue.load_state_dict(sg_sd)
    m.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
            geometric_verification=GeometricVerification.NONE)
    sg_all = (m.mkpts0.copy(), m.mkpts1.copy(), m.scores0.copy(), m.mconf.copy())
    m.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
            geometric_verification=GeometricVerification.MAGSAC)
    sg_gv = (m.mkpts0.copy(), m.mkpts1.copy())
    m.match(i0, i1, quality=Quality.MEDIUM, tile_selection=TileSelection.GRID, grid=[1, 1], overlap=0,
            geometric_verification=GeometricVerification.NONE)
    sg_med = (m.mkpts0.copy(), m.mkpts1.copy())

    # LightGlueMatcher re-instantiates both nets on every call (matchers.py:1256-1258): patch the constructors
    import icepy4d.thirdparty.LightGlue.lightglue as LGpkg
    o_sp, o_lg = LGpkg.SuperPoint, LGpkg.LightGlue
    LGpkg.SuperPoint = lambda **kw: ref_shims.build_reference_superpoint_lg(sp_sd, **kw)
    LGpkg.LightGlue = lambda features="superpoint", **kw: ref_shims.build_reference_lightglue(lg_sd, **kw)
    try:
        lm = M.LightGlueMatcher({"features": "superpoint", "force_cpu": True})
        lm.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
                 max_keypoints=512, geometric_verification=GeometricVerification.NONE)
        lg_all = (lm.mkpts0.copy(), lm.mkpts1.copy(), lm.mconf.copy())
    finally:
        LGpkg.SuperPoint, LGpkg.LightGlue = o_sp, o_lg
    np.savez_compressed(os.path.join(OUT, "matchers.npz"), image0=i0, image1=i1,
                        sg_mkpts0=sg_all[0], sg_mkpts1=sg_all[1], sg_scores0=sg_all[2], sg_mconf=sg_all[3],
                        sg_gv_mkpts0=sg_gv[0], sg_gv_mkpts1=sg_gv[1],
                        sg_med_mkpts0=sg_med[0], sg_med_mkpts1=sg_med[1],
                        lg_mkpts0=lg_all[0], lg_mkpts1=lg_all[1], lg_mconf=lg_all[2])
    print("matchers: SG", len(sg_all[0]), "SG+GV", len(sg_gv[0]), "SG medium", len(sg_med[0]), "LG", len(lg_all[0]))


def golden_tiler_quality():
    ref_shims.install_shims()
    import cv2
    from icepy4d.matching.tiling import Tiler

    rows = []
    for (h, w, grid, ov, org) in [(4000, 6000, [2, 3], 0, [0, 0]), (4000, 6000, [3, 4], 0, [0, 0]),
                                  (4008, 6012, [3, 2], 200, [0, 0]), (1000, 1500, [1, 1], 0, [0, 0]),
                                  (480, 640, [2, 2], 40, [0, 0]), (4000, 6000, [2, 3], 100, [50, 30]),
                                  (2250, 3350, [2, 2], 0, [0, 0])]:
        lims, _ = Tiler(grid=grid, overlap=ov, origin=org).compute_limits_by_grid(np.zeros((h, w), np.uint8))
        for k in sorted(lims):
            rows.append([h, w, grid[0], grid[1], ov, org[0], org[1], int(k), *[int(v) for v in lims[k]]])
    img = synthetic.blurred_noise(123, 201, 5)
    img3 = np.repeat(img[:, :, None], 3, 2)
    np.savez_compressed(os.path.join(OUT, "tiler_quality.npz"), tiler=np.array(rows, dtype=np.int64), img=img,
                        pyrdown=cv2.pyrDown(img3), pyrdown2=cv2.pyrDown(cv2.pyrDown(img3)), pyrup=cv2.pyrUp(img3),
                        gray=cv2.cvtColor(img3, cv2.COLOR_RGB2GRAY))
    print("tiler_quality: rows", len(rows))


def golden_geometry():
    ref_shims.install_shims()
    from icepy4d.matching import GeometricVerification, geometric_verification
    from icepy4d.sfm import Triangulate

    sc = synthetic.two_view_scene(n=3000, seed=7)
    cams = sc["cams"]
    F, mask = geometric_verification(sc["pts0"], sc["pts1"], method=GeometricVerification.MAGSAC)
    inl = sc["inlier"]
    tri = Triangulate(cams, [sc["pts0"][inl][:600], sc["pts1"][inl][:600]])
    X_it = tri.triangulate_two_views().copy()
    from icepy4d.sfm.geometry import undistort_points
    from icepy4d.thirdparty.triangulation import iterative_LS_triangulation
    u0 = undistort_points(sc["pts0"][inl][:600], cams[0])
    u1 = undistort_points(sc["pts1"][inl][:600], cams[1])
    _, status = iterative_LS_triangulation(u0, cams[0].P, u1, cams[1].P)
    X_lin = Triangulate(cams, [sc["pts0"][inl][:600], sc["pts1"][inl][:600]]).triangulate_two_views(
        approach="linear_triangulation").copy()
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), pts0=sc["pts0"], pts1=sc["pts1"], inlier=inl, X=sc["X"],
                        F=F, mask=mask, und0=u0, und1=u1, X_iter=X_it, status=status, X_lin=X_lin,
                        P0=cams[0].P, P1=cams[1].P)
    print("geometry: MAGSAC inliers", int(mask.sum()), "of", len(mask), "true inliers", int(inl.sum()),
          "tri rel err", float(np.median(np.linalg.norm(X_it - sc["X"][inl][:600], axis=1) / np.linalg.norm(sc["X"][inl][:600], axis=1))))


def golden_pose():
    """sfm/geometry.py:31-76 run from the reference itself on an ideal-pinhole two-view scene with 30 % outliers."""
    ref_shims.install_shims()
    from icepy4d.sfm.geometry import estimate_pose

    sc = synthetic.two_view_scene(n=4000, seed=31, noise_px=0.3, outlier_frac=0.3, distortion=False)
    cams = sc["cams"]
    R, t, inl = estimate_pose(sc["pts0"].astype(np.float64), sc["pts1"].astype(np.float64), cams[0].K, cams[1].K, 1.0, 0.9999)
    np.savez_compressed(os.path.join(OUT, "pose.npz"), pts0=sc["pts0"], pts1=sc["pts1"], inlier=sc["inlier"], K0=cams[0].K,
                        K1=cams[1].K, R_true=cams[1].R, t_true=cams[1].t, R=R, t=t, mask=inl)
    ang = np.degrees(np.arccos(np.clip((np.trace(R @ cams[1].R.T) - 1) / 2, -1, 1)))
    print("pose: inliers", int(inl.sum()), "of", len(inl), "true", int(sc["inlier"].sum()), "rotation error deg", float(ang))


def golden_colors():
    """sfm/interpolate_colors.py:14-51 run from the reference itself (its Camera class is not needed: a SimpleCamera has the
    R, t, K, dist attributes project_points reads)."""
    ref_shims.install_shims()
    from icepy4d.sfm.interpolate_colors import interpolate_point_colors

    sc = synthetic.two_view_scene(n=1500, seed=13, outlier_frac=0.0)
    cam = sc["cams"][1]
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (4008, 6012, 3), dtype=np.uint8)
    X = sc["X"].copy()
    X[:40] *= 3.0                                  # some points project outside the image: exercises the clipped weights
    col = interpolate_point_colors(X, img, cam, convert_BRG2RGB=True)
    # keep the fixture small: only the image rows/cols the points touch are needed -> store the image seed instead
    np.savez_compressed(os.path.join(OUT, "colors.npz"), X=X, R=cam.R, t=cam.t, K=cam.K, dist=cam.dist, img_seed=5,
                        img_shape=np.array(img.shape), colors=col)
    print("colors: mean", col.mean(0))


def golden_preselection():
    """TileSelection.PRESELECTION (matchers.py:513-560) and the `_match_images` / `_match_by_tile` return contracts
    (matchers.py:892-940, 394-469) run from the reference itself.  The pair is shifted by (400, 96) px so that off-diagonal tile
    pairs are selected; the per-pair counts of the pre-match are stored so that the test can check the decision margin."""
    ref_shims.install_shims()
    import cv2
    import icepy4d.matching.matchers as M
    from icepy4d.matching import GeometricVerification, Quality, TileSelection

    sp_sd, sg_sd = weights.make_superpoint_state(1), weights.make_superglue_state(2)
    i0, i1 = synthetic.stereo_pair(960, 1280, seed=1013, shift=(400, 96), channels=3)
    with ref_shims.no_checkpoint_loading():
        m = M.SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": 2048,
                                "match_threshold": 0.2, "force_cpu": True, "sinkhorn_iterations": 20})
    m.matcher.superpoint.load_state_dict(sp_sd)
    m.matcher.superglue.load_state_dict(sg_sd)
    rec = {}
    orig = m._tile_selection

    def spy(image0, image1, t0_lims, t1_lims, method=TileSelection.PRESELECTION, **config):
        pairs = orig(image0, image1, t0_lims, t1_lims, method, **config)
        if "pairs" not in rec:
            rec["pairs"], rec["t0"], rec["t1"] = list(pairs), t0_lims, t1_lims
        return pairs
    m._tile_selection = spy
    m.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.PRESELECTION, grid=[2, 3], overlap=0,
            geometric_verification=GeometricVerification.NONE)
    mk0, mk1 = m.mkpts0.copy(), m.mkpts1.copy()
    # the pre-match itself (what _tile_selection ran internally), to store the per-pair counts
    f0, f1, mtc, mconf = m._match_images(cv2.pyrDown(i0), cv2.pyrDown(i1), max_keypoints=4096)
    vld = mtc > -1
    kp0, kp1 = f0.keypoints[vld] * 2, f1.keypoints[mtc[vld]] * 2
    keys0, keys1 = sorted(rec["t0"]), sorted(rec["t1"])
    counts = np.zeros((len(keys0), len(keys1)), np.int64)
    inr = lambda p, r: np.all(p > r[:2], axis=1) & np.all(p < r[2:], axis=1)
    for a in keys0:
        for b in keys1:
            counts[a, b] = int(np.sum(inr(kp0, rec["t0"][a]) & inr(kp1, rec["t1"][b])))
    assert sorted(rec["pairs"]) == sorted((a, b) for a in keys0 for b in keys1 if counts[a, b] > 5)
    # _match_images / _match_by_tile return contracts on a small pair
    j0, j1 = synthetic.stereo_pair(240, 320, seed=1012, shift=(16, 8), channels=3)
    g0, g1, gm, gc = m._match_images(j0, j1)
    t0, t1, tm, tc = m._match_by_tile(j0, j1, tile_selection=TileSelection.GRID, grid=[1, 2], overlap=20)
    # images are regenerated from their seeds by the test (synthetic.stereo_pair is deterministic); descriptors: first 64 columns
    np.savez_compressed(os.path.join(OUT, "preselection.npz"), seed=1013, shift=np.array([400, 96]), shape=np.array([960, 1280]),
                        pairs=np.array(sorted(rec["pairs"]), np.int64), counts=counts, mkpts0=mk0, mkpts1=mk1,
                        small_seed=1012, small_shift=np.array([16, 8]), small_shape=np.array([240, 320]),
                        mi_kpts0=g0.keypoints, mi_desc0=g0.descriptors[:, :64], mi_desc_shape=np.array(g0.descriptors.shape),
                        mi_scores0=g0.scores, mi_kpts1=g1.keypoints, mi_scores1=g1.scores, mi_matches0=gm, mi_mconf=gc,
                        mt_kpts0=t0.keypoints, mt_kpts1=t1.keypoints, mt_desc0=t0.descriptors[:, :64],
                        mt_desc_shape=np.array(t0.descriptors.shape), mt_scores0=t0.scores, mt_matches0=tm, mt_mconf=tc)
    print("preselection: pairs", sorted(rec["pairs"]), "counts", counts.tolist(), "matches", len(mk0),
          "| _match_images", g0.keypoints.shape, g0.descriptors.shape, gm.shape, gc.shape, gm.dtype,
          "| _match_by_tile", t0.keypoints.shape, t0.descriptors.shape, tm.shape, tc.shape)


def golden_helmert():
    """sfm/absolute_orientation.py:56-287 run from the reference itself: closed-form similarity between two noisy point sets,
    its Euler / translation parameters, and the transformation applied to a point cloud and to two cameras."""
    ref_shims.install_shims()
    from icepy4d.sfm.absolute_orientation import Absolute_orientation

    rng = np.random.default_rng(21)
    n = 12
    v0 = rng.uniform(-50, 50, (n, 3))
    ang = np.array([0.3, -0.2, 1.1])
    cx, cy, cz = np.cos(ang); sx, sy, sz = np.sin(ang)
    R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    v1 = 2.5 * v0 @ R.T + np.array([1000.0, -2000.0, 300.0]) + rng.normal(0, 0.05, (n, 3))
    cams = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    ext0 = [c.extrinsics.copy() for c in cams]
    ao = Absolute_orientation(tuple(cams), points3d_final=v1.copy(), points3d_orig=v0.copy())
    T = ao.estimate_transformation_linear()
    T_noscale = Absolute_orientation((), points3d_final=v1.copy(), points3d_orig=v0.copy()).estimate_transformation_linear(estimate_scale=False)
    prm = ao.extract_params_from_T()
    cloud = rng.uniform(-80, 80, (5000, 3))
    out = ao.apply_transformation(points3d=cloud.copy())
    ext1 = [c.extrinsics.copy() for c in cams]
    np.savez_compressed(os.path.join(OUT, "helmert.npz"), v0=v0, v1=v1, T=T, T_noscale=T_noscale,
                        params=np.array([prm[k] for k in ("rx", "ry", "rz", "tx", "ty", "tz", "m")]), cloud=cloud, cloud_out=out,
                        ext_before=np.array(ext0), ext_after=np.array(ext1))
    print("helmert: scale", np.cbrt(np.linalg.det(T[:3, :3])), "params", prm)


def golden_containers():
    """core/features.py:208-632, core/points.py:172-514 and io/export2bundler.py:90-172 run from the reference itself: the
    Features / Points fill (`append_*_from_numpy`), their read-out and filters, and the Bundler .out file for a two-camera epoch."""
    ref_shims.install_shims()
    import tempfile
    from icepy4d.core import Camera, Features, Points
    from icepy4d.io.export2bundler import write_bundler_out

    rng = np.random.default_rng(33)
    n = 240
    x, y = rng.uniform(1, 6000, n).astype(np.float32), rng.uniform(1, 4000, n).astype(np.float32)
    descr = rng.normal(size=(128, n)).astype(np.float32)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    out = {"x": x, "y": y, "descr": descr, "scores": scores}
    f = Features()
    f.append_features_from_numpy(x[:160], y[:160], descr[:, :160], scores[:160], epoch=3)
    ids2 = [int(i) for i in range(1000, 1080)]
    f.append_features_from_numpy(x[160:], y[160:], descr[:, 160:], scores[160:], track_ids=ids2, epoch=3)
    out["f_len"], out["f_last"] = np.array(len(f)), np.array(f.last_track_id)
    out["f_ids"] = np.array(f.get_track_ids())
    d = f.to_numpy(get_descr=True, get_score=True)
    out["f_kpts"], out["f_descr"], out["f_scores"] = d["kpts"], d["descr"], d["scores"]
    one = f[1005]
    out["f_1005"] = np.concatenate([one.xy.reshape(-1), [one.score], one.descr.reshape(-1)[:4], [one.track_id], [one.epoch]]).astype(np.float64)
    out["f_1005_descr_shape"] = np.array(one.descr.shape)
    # duplicate ids -> progressive ids
    f.append_features_from_numpy(x[:5], y[:5], descr[:, :5], scores[:5], track_ids=[0, 1, 2, 3, 4])
    out["f_ids_dup"] = np.array(f.get_track_ids())
    mask = rng.uniform(size=len(f)) < 0.6
    f.filter_feature_by_mask(list(mask))
    out["mask"], out["f_ids_masked"], out["f_kpts_masked"] = mask, np.array(f.get_track_ids()), f.kpts_to_numpy()
    keep = [int(i) for i in np.array(f.get_track_ids())[::3]]
    f.filter_feature_by_index(keep)
    out["f_ids_indexed"], out["f_scores_indexed"] = np.array(f.get_track_ids()), f.scores_to_numpy()
    out["f_contains"] = np.array([keep[0] in f, 999999 in f])

    m = 200
    fa, fb = Features(), Features()
    fa.append_features_from_numpy(x[:m], y[:m], descr[:, :m], scores[:m])
    fb.append_features_from_numpy(x[:m] - rng.uniform(5, 40, m).astype(np.float32), y[:m] + rng.normal(0, 1, m).astype(np.float32),
                                  descr[:, :m], scores[:m])
    out["fb_kpts"] = fb.kpts_to_numpy()
    xyz = rng.uniform(-300, 300, (m, 3))
    colors = rng.uniform(0, 1, (m, 3))
    pts = Points()
    pts.append_points_from_numpy(xyz[:120], colors=colors[:120])
    pts.append_points_from_numpy(xyz[120:], track_ids=[int(i) for i in range(120, m)], colors=colors[120:])
    out["xyz"], out["colors"] = xyz, colors
    out["p_xyz"], out["p_col"], out["p_col8"] = pts.to_numpy(), pts.colors_to_numpy(), pts.colors_to_numpy(as_uint8=True)
    out["p_ids"], out["p_last"] = np.array(pts.get_track_ids()), np.array(pts.last_track_id)
    sc = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    cams = {"cam1": Camera(6012, 4008, K=sc[0].K.copy(), dist=sc[0].dist.copy(), R=sc[0].R.copy(), t=sc[0].t.copy()),
            "cam2": Camera(6012, 4008, K=sc[1].K.copy(), dist=sc[1].dist.copy(), R=sc[1].R.copy(), t=sc[1].t.copy())}

    class _Img:
        def __init__(self, path):
            self.path = path
    with tempfile.TemporaryDirectory() as td:
        write_bundler_out(td, "epoch", {"cam1": _Img("/data/cam1/a.jpg"), "cam2": _Img("/data/cam2/a.jpg")}, cams,
                          {"cam1": fa, "cam2": fb}, pts)
        out["bundler_out"] = np.frombuffer(open(os.path.join(td, "epoch.out"), "rb").read(), dtype=np.uint8)
        out["bundler_imlist"] = np.frombuffer(open(os.path.join(td, "im_list.txt"), "rb").read(), dtype=np.uint8)
    pmask = rng.uniform(size=m) < 0.5
    pts.filter_point_by_mask(pmask)
    out["pmask"], out["p_ids_masked"], out["p_last_masked"] = pmask, np.array(pts.get_track_ids()), np.array(pts.last_track_id)
    np.savez_compressed(os.path.join(OUT, "containers.npz"), **out)
    print("containers: features", out["f_len"], "points", len(out["p_xyz"]), "bundler bytes", len(out["bundler_out"]))


if __name__ == "__main__":
    assert ref_shims.reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:                       # regenerate selected fixtures only: python -m oracle.make_golden preselection ...
        for name in sys.argv[1:]:
            globals()[f"golden_{name}"]()
        sys.exit(0)
    golden_tiler_quality()
    golden_geometry()
    golden_pose()
    golden_colors()
    golden_superpoint_superglue()
    golden_lightglue()
    golden_matchers()
    golden_preselection()
    golden_helmert()
    golden_containers()
