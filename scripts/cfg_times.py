"""Informational timings (CUDA events) of the non-headline BASELINE configs on one GPU: cfg1, cfg3, cfg5."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icepy4d_b200 import ops, synthetic, weights
from icepy4d_b200.epoch import make_cfg5_pipeline
from icepy4d_b200.matching import GeometricVerification, LightGlueMatcher, Quality, TileSelection
from icepy4d_b200.sfm import Triangulate


def ev(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


which = sys.argv[1:] or ["cfg1", "cfg3", "cfg5"]
if "cfg1" in which:
    i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1003, shift=(32, 16), channels=3)
    m = LightGlueMatcher({"features": "superpoint", "superpoint_state": weights.make_superpoint_state(1),
                          "lightglue_state": weights.make_lightglue_state(3), "precision": "bf16", "conv_precision": "f16x3"})
    f = lambda: m.match(i0, i1, quality=Quality.LOW, tile_selection=TileSelection.GRID, grid=[1, 1], overlap=0, max_keypoints=2048,
                        geometric_verification=GeometricVerification.MAGSAC)
    t0 = time.perf_counter(); f(); torch.cuda.synchronize()
    ms = ev(f)
    print(f"cfg1 (6000x4000 -> Quality.LOW 1500x1000, LightGlue 2048 kp, whole match() incl. H2D/pyrDown/GV): {ms:.1f} ms, {len(m.mkpts0)} inliers")
if "cfg3" in which:
    sc = synthetic.two_view_scene(n=200_000, seed=11)
    p0, p1 = torch.from_numpy(sc["pts0"]).cuda(), torch.from_numpy(sc["pts1"]).cuda()
    ms = ev(lambda: ops.fundamental_ransac(p0, p1, 0.5, 0.999, 100000, seed=0), reps=5)
    print(f"cfg3 F-matrix RANSAC + polish, 200k correspondences: {ms:.2f} ms ({200e3 / ms / 1e3:.1f} M correspondences/s)")
    inl = torch.from_numpy(sc["inlier"]).cuda()
    q0, q1 = p0[inl].contiguous(), p1[inl].contiguous()
    cams = sc["cams"]
    def tri():
        u0 = ops.undistort_points(q0, cams[0].K, cams[0].dist); u1 = ops.undistort_points(q1, cams[1].K, cams[1].dist)
        return ops.triangulate_iterative_ls(u0, u1, cams[0].P, cams[1].P)
    ms = ev(tri, reps=10)
    print(f"cfg3 undistort x2 + iterative-LS triangulation, {q0.shape[0]} points: {ms * 1e3:.1f} us ({q0.shape[0] / ms / 1e3:.1f} M points/s)")
    def dlt():
        u0 = ops.undistort_points(q0, cams[0].K, cams[0].dist); u1 = ops.undistort_points(q1, cams[1].K, cams[1].dist)
        return ops.triangulate_dlt(u0, u1, cams[0].P, cams[1].P)
    ms = ev(dlt, reps=10)
    print(f"cfg3 undistort x2 + DLT (6x6 Jacobi SVD), {q0.shape[0]} points: {ms * 1e3:.1f} us ({q0.shape[0] / ms / 1e3:.1f} M points/s)")
if "cfg5" in which:
    pipe = make_cfg5_pipeline(16384, precision="bf16", conv_precision="f16x3", grid=(3, 4))
    i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1005, shift=(16, 8), channels=3)
    d0, d1 = torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()
    ms = ev(lambda: pipe.run_device(d0, d1), reps=2)
    out = pipe.run_device(d0, d1)
    print(f"cfg5 (6000x4000, 3x4 tiles, LightGlue 16384 kp/tile, static depth/width, dual-softmax + mutual NN): {ms:.1f} ms/epoch "
          f"= {1e3 / ms:.2f} epochs/s on one GPU, {out['mkpts0'].shape[0]} verified matches")
