"""Per-layer timings (CUDA events) of the split-bf16 tcgen05 SuperPoint backbone at one 2000x2000 tile, next to cuDNN TF32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops, weights
from icepy4d_b200.matching.superpoint import SuperPointB200

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 2000


def ev(fn, reps=5, warm=2):
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


st = weights.make_superpoint_state(1)
sp = SuperPointB200(st, conv_precision="bf16x3")
img = torch.rand(1, 1, H, W, device="cuda")
pk = sp.pk
x1 = ops.sp_conv1a_relu_split(img, *sp.w1a)
t = ev(lambda: ops.sp_conv1a_relu_split(img, *sp.w1a))
print(f"conv1a (SIMT f32 -> split planes)        {t * 1e3:8.1f} us   {H * W * 64 * 4 / t / 1e6:7.0f} GB/s written")
t = ev(lambda: ops.sp_conv1ab_fused(img, sp.w1a[0], sp.w1a[1], sp.pk["conv1b"], pool=True))
print(f"conv1a+conv1b fused (one kernel, pooled)  {t * 1e3:8.1f} us")
plan = [("conv1b", True), ("conv2a", False), ("conv2b", True), ("conv3a", False), ("conv3b", True), ("conv4a", False), ("conv4b", False)]
x = x1
total = t
for name, pool in plan:
    p = pk[name]
    h, w = x.shape[1], x.shape[2]
    fl = 2.0 * h * w * p.cin * p.cout * 9
    t = ev(lambda: ops.conv_bf16x3(x, p, pool=pool))
    total += t
    print(f"{name} {p.cin:3d}->{p.cout:3d} @{h}x{w} pool={int(pool)}        {t * 1e3:8.1f} us   {fl / t / 1e9:7.1f} TFLOP/s useful, {3 * fl / t / 1e9:7.1f} TFLOP/s bf16 MMA")
    x = ops.conv_bf16x3(x, p, pool=pool)
for a, b, out in (("convPa", "convPb", "planar"), ("convDa", "convDb", "nhwc")):
    h, w = x.shape[1], x.shape[2]
    p = pk[a]
    fl = 2.0 * h * w * p.cin * p.cout * 9
    t = ev(lambda: ops.conv_bf16x3(x, p))
    total += t
    print(f"{a} {p.cin:3d}->{p.cout:3d} @{h}x{w}               {t * 1e3:8.1f} us   {fl / t / 1e9:7.1f} TFLOP/s useful, {3 * fl / t / 1e9:7.1f} TFLOP/s bf16 MMA")
    y = ops.conv_bf16x3(x, p)
    p = pk[b]
    fl = 2.0 * h * w * p.cin * p.cout
    t = ev(lambda: ops.conv_bf16x3(y, p, relu=False, out=out))
    total += t
    print(f"{b} {p.cin:3d}->{p.cout:3d} @{h}x{w} (1x1)         {t * 1e3:8.1f} us   {fl / t / 1e9:7.1f} TFLOP/s useful")
print(f"sum of layers                             {total * 1e3:8.1f} us")
t = ev(lambda: sp.backbone(img))
print(f"backbone bf16x3 (whole)                   {t * 1e3:8.1f} us   {676.4 / t:.1f} TFLOP/s useful")
sp2 = SuperPointB200(st, conv_precision="tf32")
t = ev(lambda: sp2.backbone(img))
print(f"backbone cuDNN tf32 (whole)               {t * 1e3:8.1f} us   {676.4 / t:.1f} TFLOP/s")
