"""One sixth of a cfg2 epoch (2 SuperPoint tiles + 1 SuperGlue pair at 8192 kp, 100 Sinkhorn iterations) for an ncu launch
list: the epoch is 6 identical slices, so kernel SHARES are those of the full step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import synthetic
from icepy4d_b200.epoch import make_cfg2_pipeline
from icepy4d_b200.matching import GeometricVerification, Quality, TileSelection
pipe = make_cfg2_pipeline(8192, 100, precision="bf16", conv_precision="f16x3", grid=(1, 1))
i0, i1 = synthetic.stereo_pair(1999, 1999, seed=1000, shift=(16, 24), channels=3)
d0, d1 = torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()
out = pipe.run_device(d0, d1)            # warm-up (cuDNN autotune, allocator), not profiled
torch.cuda.synchronize()
torch.cuda.profiler.start()              # ncu --profile-from-start off
out = pipe.run_device(d0, d1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("matches", out["mkpts0"].shape[0])
