"""Where does the layer-GEMM time go?  GPU time (ncu-free: CUDA events over 200 back-to-back launches on pre-encoded problem
sizes) as a function of K (operand traffic / MMA) and N (tiles, epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops_tc
M = 16384
def t(N, K, reps=200, **kw):
    x = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops_tc.gemm_tc(x, W, b, out16=o, relu=True)
    for _ in range(5): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 20): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
NS = (768,) if os.environ.get("I4D_GEMM_DBG") else (128, 256, 512, 768)
for N in NS:
    print(f"N={N:4d}: " + "  ".join(f"K={K}: {t(N, K):6.1f} us" for K in ((64, 256) if os.environ.get("I4D_GEMM_DBG") else (64, 128, 256, 512, 1024))))
# cuBLAS (torch.mm, bf16) on the same shapes, for orientation only (library code is not on the product path)
def tl(N, K, reps=200):
    x = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    f = lambda: torch.mm(x, W.t(), out=o)
    for _ in range(5): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 20): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
if not os.environ.get("I4D_GEMM_DBG"):
    for N in (256, 512, 768):
        print(f"cuBLAS N={N:4d}: " + "  ".join(f"K={K}: {tl(N, K):6.1f} us" for K in (256, 512)))
