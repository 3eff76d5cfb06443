"""Fused Sinkhorn at cfg2 size (8192^2, 100 iterations): us/iteration with CUDA events (not under a profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
M = int(sys.argv[3]) if len(sys.argv) > 3 else N          # rows (keypoints of image 0); N = columns
S = ops.padded_scores(M, N, "cuda")
S.copy_(torch.randn(M, N, device="cuda") * 3)
ws = ops.AssignWorkspace(M, N, S.device)
for mode in (0, 2):
    ops.set_sinkhorn_mode(mode)
    for _ in range(2):
        ops.sinkhorn(S, 1.0, iters, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        ops.sinkhorn(S, 1.0, iters, ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{M}x{N} mode {mode}: {ms:.3f} ms / {iters} iterations = {ms / iters * 1e3:.2f} us/iter, "
          f"{2 * M * N * 4 * iters / ms / 1e6:.0f} GB/s algorithmic, {M * N * 4 * iters / ms / 1e6:.0f} GB/s single-read")
ops.set_sinkhorn_mode(0)
