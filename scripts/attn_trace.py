"""Cycle timeline of attn_tc_kernel: builds a second library with -DFA_TRACE (the product library is untouched), runs the cfg2-size
launch and prints, for CTA 0, the %clock64 stamps of one softmax warp per tile (steady-state blocks)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icepy4d_b200 import build as B, _native
obj = "/tmp/attn_tc_trace.o"; lib = "/tmp/libicepy4d_trace.so"
subprocess.check_call([B.NVCC, *B.ARCH, *[c for c in B.COMMON if c not in ("-Xptxas", "-v")], "-w", "-DFA_TRACE", "-c", os.path.join(B.CSRC, "attn_tc.cu"), "-o", obj])
objs = [os.path.join(B.OUT_DIR, f) for f in os.listdir(B.OUT_DIR) if f.endswith(".o") and f != "attn_tc.o"]
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-cudart", "static"])
_native.LIB_PATH = lib
import numpy as np
import torch
from icepy4d_b200 import ops_tc
N = int(os.environ.get("ATTN_N", "8192"))
qkv = (torch.randn(2 * N, 768, device="cuda") * 1.5).bfloat16()
att = torch.empty(2 * N, 256, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops_tc.attention_tc(qkv, [(0, N, 0, N), (N, N, N, N)], att, 0, 256, 512)
torch.cuda.synchronize()
buf = np.zeros((2, 96, 8), dtype=np.uint64)
assert _native.lib().i4d_attention_trace_dump(ctypes.c_void_p(buf.ctypes.data)) == 0
rel = buf.astype(np.int64) - int(buf[buf > 0].min())
cols = ["loop top", "S ready", "max done", "max exchanged", "PV(j-1) retired", "exps done", "P handed over"]
print(f"columns: {cols}")
for a, name in enumerate(("softmax A", "softmax B")):
    print(f"== {name}: clk since the first stamp, blocks 20..25")
    for b in range(20, 26):
        print(f"  blk {b}: " + " ".join(f"{int(x):7d}" for x in rel[a, b, :7]))
    print(f"  -> {(rel[a, 60, 1] - rel[a, 20, 1]) / 40.0:.0f} clk per block; mean deltas: " + " ".join(f"{x:6.0f}" for x in np.diff(rel[a, 20:60, :7], axis=1).mean(0)))
print("B minus A at 'S ready' (blocks 20..60): mean %.0f clk" % (rel[1, 20:60, 1] - rel[0, 20:60, 1]).mean())
