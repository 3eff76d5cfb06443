"""Builds attn_tc.cu variants with -D flags into separate libraries and times each at 8192^2 and 16384^2."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icepy4d_b200 import build as B
flags = sys.argv[1].split() if len(sys.argv) > 1 and sys.argv[1] else []
obj = "/tmp/attn_var.o"; lib = "/tmp/libicepy4d_var.so"
subprocess.check_call([B.NVCC, *B.ARCH, *[c for c in B.COMMON if c not in ("-Xptxas", "-v")], "-w", *flags, "-c", os.path.join(B.CSRC, "attn_tc.cu"), "-o", obj])
objs = [os.path.join(B.OUT_DIR, f) for f in os.listdir(B.OUT_DIR) if f.endswith(".o") and f != "attn_tc.o"]
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-cudart", "static"])
from icepy4d_b200 import _native
_native.LIB_PATH = lib
import torch
from icepy4d_b200 import ops_tc
for N in (8192, 16384):
    qkv = (torch.randn(2 * N, 768, device="cuda") * 1.5).bfloat16(); att = torch.empty(2 * N, 256, device="cuda", dtype=torch.bfloat16)
    pr = [(0, N, 0, N), (N, N, N, N)]
    for _ in range(3): ops_tc.attention_tc(qkv, pr, att, 0, 256, 512)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops_tc.attention_tc(qkv, pr, att, 0, 256, 512)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    q = qkv[:512, :64].double(); k = qkv[:N, 256:320].double(); v = qkv[:N, 512:576].double()
    err = (att[:512, :64].double() - torch.softmax(q @ k.t() * 0.125, -1) @ v).abs().max().item()
    print(f"flags={' '.join(flags) or '-'} POLY={os.environ.get('I4D_FA_POLY','default')} N={N}: {ms*1e3:.1f} us  {2*4*(2*N*N*64*2)/ms/1e9:.0f} TFLOP/s  max abs err {err:.4f}", flush=True)
