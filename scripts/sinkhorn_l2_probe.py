"""How fast is the fused Sinkhorn band loop when the score matrix is L2-resident?  Times M x 8192 problems (M*8192*4 bytes:
33 MB ... 268 MB) for 100 iterations; run once plainly and once with I4D_SK_DBG=7 (barriers and combine off: band loop only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops

N, iters = 8192, 100
for M in (1024, 2048, 3072, 4096, 6144, 8192):
    S = torch.randn(M, N, device="cuda") * 3
    ws = ops.AssignWorkspace(M, N, S.device)
    for _ in range(2):
        ops.sinkhorn(S, 1.0, iters, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.sinkhorn(S, 1.0, iters, ws)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 3 / iters * 1e3
    print(f"I4D_SK_DBG={os.environ.get('I4D_SK_DBG', '0')}  M={M:5d}  S={M * N * 4 / 1e6:6.1f} MB  {us:6.2f} us/iter  "
          f"{us / M * 1e3:6.2f} ns/row  {M * N * 4 / us / 1e3:7.1f} GB/s single-read", flush=True)
