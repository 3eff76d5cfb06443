"""Is the ~1.8 ms device gap the profiler shows between the first tile_gray_kernel and the first convolution of an epoch real?
CUDA events (no profiler) right after the first grey conversion and right before / after the first backbone kernel, plus the
host clock at the same places."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops, synthetic
from icepy4d_b200.epoch import make_cfg2_pipeline

pipe = make_cfg2_pipeline(8192, 100, precision="bf16", conv_precision="f16x3")
i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1000, shift=(16, 24), channels=3)
d0, d1 = torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()
for _ in range(3):
    pipe.run_device(d0, d1)
torch.cuda.synchronize()
gray, conv = ops.tile_to_gray_f32, ops.sp_conv1ab_fused
marks = []


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def gray_spy(*a, **k):
    out = gray(*a, **k)
    if len(marks) < 8:
        marks.append(("after tile_gray", ev(), time.perf_counter()))
    return out


def conv_spy(*a, **k):
    if len(marks) < 8:
        marks.append(("before conv1ab", ev(), time.perf_counter()))
    out = conv(*a, **k)
    if len(marks) < 8:
        marks.append(("after conv1ab", ev(), time.perf_counter()))
    return out


ops.tile_to_gray_f32, ops.sp_conv1ab_fused = gray_spy, conv_spy
for rep in range(3):
    marks.clear()
    e0, t0 = ev(), time.perf_counter()
    pipe.run_device(d0, d1)
    e1 = ev()
    torch.cuda.synchronize()
    print(f"epoch {rep}: GPU span {e0.elapsed_time(e1):.2f} ms")
    for name, e, t in marks:
        print(f"   {name:18s} device +{e0.elapsed_time(e):7.3f} ms   host +{1e3 * (t - t0):7.3f} ms")
