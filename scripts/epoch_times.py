"""Per-epoch device time of the cfg2 pipeline (CUDA events around every run_device call): shows warm-up effects and outliers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import _native, synthetic
from icepy4d_b200.epoch import make_cfg2_pipeline

pipe = make_cfg2_pipeline(8192, 100, precision="bf16", conv_precision="f16x3", grid=(2, 3))
dev = []
for e in range(2):
    i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1000 + e, shift=(16, 24), channels=3)
    dev.append((torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()))
keep = os.environ.get("KEEP", "0") == "1"
res = []
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _native.LAUNCHES
    e0.record()
    r = pipe.run_device(*dev[s % 2])
    e1.record()
    torch.cuda.synchronize()
    if keep:
        res.append(r["points3d"])
    print(f"epoch {s:2d}: {e0.elapsed_time(e1):7.2f} ms  launches {_native.LAUNCHES - n0}  matches {r['mkpts0'].shape[0]}  "
          f"mem {torch.cuda.memory_allocated() / 1e9:.2f} GB reserved {torch.cuda.memory_reserved() / 1e9:.2f} GB", flush=True)
