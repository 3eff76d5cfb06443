"""Attention kernel at cfg2 size (2 images x 4 heads x 8192^2, head_dim 64): CUDA-event timing, max error vs an f64 reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops_tc
N = 8192
nt = 2 * N
torch.manual_seed(0)
qkv = (torch.randn(nt, 768, device="cuda") * 1.5).bfloat16()
att = torch.empty(nt, 256, device="cuda", dtype=torch.bfloat16)
probs = [(0, N, 0, N), (N, N, N, N)]
for _ in range(3):
    ops_tc.attention_tc(qkv, probs, att, 0, 256, 512)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops_tc.attention_tc(qkv, probs, att, 0, 256, 512)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 2 * 4 * (2 * N * N * 64 * 2)
# accuracy on the first 512 queries of image 0, head 0
q = qkv[:512, :64].double(); k = qkv[:N, 256:320].double(); v = qkv[:N, 512:576].double()
ref = torch.softmax(q @ k.t() * 0.125, -1) @ v
err = (att[:512, :64].double() - ref).abs().max().item()
print(f"I4D_FA_POLY={os.environ.get('I4D_FA_POLY', 'default')}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  max abs err {err:.4f}")
