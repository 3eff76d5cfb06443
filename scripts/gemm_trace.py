"""Cycle timeline of gemm_tc_kernel: builds a second library with -DGT_TRACE (the product library is untouched), runs a layer GEMM
with many tiles per CTA and prints, for CTA 0, the %clock64 stamps of the producer lane, the MMA lane and epilogue thread 0."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icepy4d_b200 import build as B, _native
obj = "/tmp/gemm_tc_trace.o"; lib = "/tmp/libicepy4d_gtrace.so"
subprocess.check_call([B.NVCC, *B.ARCH, *[c for c in B.COMMON if c not in ("-Xptxas", "-v")], "-w", "-DGT_TRACE", "-c", os.path.join(B.CSRC, "gemm_tc.cu"), "-o", obj])
objs = [os.path.join(B.OUT_DIR, f) for f in os.listdir(B.OUT_DIR) if f.endswith(".o") and f != "gemm_tc.o"]
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-cudart", "static"])
_native.LIB_PATH = lib
import numpy as np
import torch
from icepy4d_b200 import ops_tc
M = int(os.environ.get("GEMM_M", "65536"))
for N, K in ((768, 256), (768, 64), (512, 512)):
    x = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops_tc.gemm_tc(x, W, b, out16=o, relu=True)
    torch.cuda.synchronize()
    buf = np.zeros((3, 32, 8), dtype=np.uint64)
    assert _native.lib().i4d_gemm_trace_dump(ctypes.c_void_p(buf.ctypes.data)) == 0
    t0 = int(buf[buf > 0].min())
    rel = buf.astype(np.int64) - t0
    tiles = (M // 128) * (N // 128) // 148
    hi = min(tiles - 1, 30)
    print(f"==== M={M} N={N} K={K}: {tiles} tiles per CTA; steady state = tiles 6..{hi}")
    names = (("producer", ["tile top", "stage 0 free", "last stage free"]),
             ("MMA lane", ["tile top", "accumulator free", "stage 0 landed", "last stage landed", "commit issued"]),
             ("epilogue thread 0", ["tile top", "top barrier", "accumulator full", "in registers", "step 0 staged", "step 0 barrier", "step 1 staged", "step 1 barrier"]))
    for r, (role, cols) in enumerate(names):
        n = len(cols)
        print(f"-- {role}: {cols}")
        for t in (6, 7, 8):
            print(f"   tile {t}: " + " ".join(f"{int(v):7d}" for v in rel[r, t, :n]))
        per = (rel[r, hi, 0] - rel[r, 6, 0]) / (hi - 6)
        d = np.diff(rel[r, 6:hi, :n], axis=1).mean(0)
        print(f"   {per:.0f} clk per tile; mean deltas between stamps: " + " ".join(f"{v:6.0f}" for v in d) + f" | to next tile top {per - d.sum():6.0f}")
