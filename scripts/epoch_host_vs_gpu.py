"""Is the cfg2 epoch bound by the GPU or by the host's launch path?  Times one epoch three ways: host enqueue only (no
sync), GPU span by CUDA events, and wall clock with a final sync."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import synthetic, _native
from icepy4d_b200.epoch import make_cfg2_pipeline

pipe = make_cfg2_pipeline(8192, 100, precision="bf16", conv_precision="f16x3")
i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1000, shift=(16, 24), channels=3)
d0, d1 = torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()
for _ in range(3):
    pipe.run_device(d0, d1)
torch.cuda.synchronize()
for rep in range(3):
    n0 = _native.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    pipe.run_device(d0, d1)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0):.1f} ms | GPU span {e0.elapsed_time(e1):.1f} ms | wall {1e3 * (t2 - t0):.1f} ms | "
          f"own launches {_native.LAUNCHES - n0}")
