"""Small driver for ncu captures of the dominant kernels at cfg2 sizes (one launch each after warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops, ops_tc
which = sys.argv[1] if len(sys.argv) > 1 else "all"
N = 8192
if which in ("sinkhorn", "all"):
    S = torch.randn(N, N, device="cuda") * 3
    ws = ops.AssignWorkspace(N, N, S.device)
    for _ in range(2):
        ops.sinkhorn(S, 1.0, 20, ws)
    torch.cuda.synchronize()
if which in ("attn", "all"):
    qkv = torch.randn(2 * N, 768, device="cuda").bfloat16()
    att = torch.empty(2 * N, 256, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops_tc.attention_tc(qkv, [(0, N, 0, N), (N, N, N, N)], att, 0, 256, 512)
    torch.cuda.synchronize()
if which in ("gemm", "all"):
    nt = 2 * N
    x = torch.randn(nt, 512, device="cuda").bfloat16()
    x32 = torch.randn(nt, 256, device="cuda")
    for (K, Nn, res) in ((256, 768, False), (512, 512, False), (512, 256, True)):
        W = torch.randn(Nn, K, device="cuda").bfloat16(); b = torch.randn(Nn, device="cuda")
        o = torch.empty(nt, Nn, device="cuda", dtype=torch.bfloat16)
        for _ in range(2):
            if res:
                ops_tc.gemm_tc(x[:, :K], W, b, residual=x32, out32=x32, out16=o)
            else:
                ops_tc.gemm_tc(x[:, :K], W, b, out16=o, relu=True)
    torch.cuda.synchronize()
if which in ("assign", "all"):
    N2 = 16384
    S2 = torch.randn(N2, N2, device="cuda") * 3
    ws2 = ops.AssignWorkspace(N2, N2, S2.device)
    z0, z1 = torch.randn(N2, device="cuda"), torch.randn(N2, device="cuda")
    for _ in range(2):
        ops.lg_assign(S2, z0, z1, 0.1, ws2)
    torch.cuda.synchronize()
