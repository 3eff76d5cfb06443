"""Cycle timeline of the fused Sinkhorn kernel's fast-mode stages: builds a second library with -DSK_TRACE (the product library is
untouched), runs 8192^2 x 12 iterations and prints, for CTA 0 (warps 0 and 9), the mean %clock64 deltas between the
synchronisation points of a stage."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icepy4d_b200 import build as B, _native
obj = "/tmp/assignment_trace.o"; lib = "/tmp/libicepy4d_sktrace.so"
subprocess.check_call([B.NVCC, *B.ARCH, *[c for c in B.COMMON if c not in ("-Xptxas", "-v")], "-w", "-DSK_TRACE", "-c", os.path.join(B.CSRC, "assignment.cu"), "-o", obj])
objs = [os.path.join(B.OUT_DIR, f) for f in os.listdir(B.OUT_DIR) if f.endswith(".o") and f != "assignment.o"]
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-cudart", "static"])
_native.LIB_PATH = lib
import numpy as np
import torch
from icepy4d_b200 import ops
N = 8192
S = torch.randn(N, N, device="cuda") * 3
ws = ops.AssignWorkspace(N, N, S.device)
for _ in range(2):
    ops.sinkhorn(S, 1.0, 12, ws)
torch.cuda.synchronize()
buf = np.zeros((2, 512, 8), dtype=np.uint64)
assert _native.lib().i4d_sinkhorn_trace_dump(ctypes.c_void_p(buf.ctypes.data)) == 0
cols = ["stage top", "rows landed (TMA)", "phase A done (exps, row partials)", "partials published", "all partials in (prev stage)", "column pass of prev stage done"]
print("columns:", cols)
for a, name in enumerate(("warp 0", "warp 9")):
    b = buf[a].astype(np.int64)
    ok = (b[:, :6] > 0).all(1)
    idx = np.flatnonzero(ok)[20:260]; idx = idx[idx < 440]
    d = np.diff(b[idx][:, :6], axis=1)
    per = np.diff(b[idx, 0])
    print(f"== {name}: {len(idx)} stages; clk per stage (top to top) median {np.median(per):.0f}, mean {per.mean():.0f}")
    print("   mean deltas between consecutive stamps:", " ".join(f"{x:7.0f}" for x in d.mean(0)), "  | medians:", " ".join(f"{x:6.0f}" for x in np.median(d, 0)))
bt = buf[1, 448:448 + 10].astype(np.int64)
bt = bt[(bt[:, :6] > 0).all(1)][2:]
d = np.diff(bt[:, :6], axis=1)
print("== iteration boundary, CTA 0 thread 0 (fast iterations): band pass | sums staged (fixed point) | bulk reduction into L2 done | grid barrier | totals read, new v")
print("   mean clk:", " ".join(f"{x:8.0f}" for x in d.mean(0)), "  -> iteration", f"{np.diff(bt[:, 0]).mean():.0f} clk")
