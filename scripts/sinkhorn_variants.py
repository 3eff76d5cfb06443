"""Builds assignment.cu with -D flags into a separate library and times / checks the fused Sinkhorn (I4D_SK_THREADS=256 / 512 forces
an instantiation):  I4D_SK_THREADS=512 python scripts/sinkhorn_variants.py "" """
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icepy4d_b200 import build as B
flags = sys.argv[1].split() if len(sys.argv) > 1 and sys.argv[1] else []
obj = "/tmp/assign_var.o"; lib = "/tmp/libicepy4d_skvar.so"
subprocess.check_call([B.NVCC, *B.ARCH, *[c for c in B.COMMON if c not in ("-Xptxas", "-v")], "-w", *flags, "-c", os.path.join(B.CSRC, "assignment.cu"), "-o", obj])
objs = [os.path.join(B.OUT_DIR, f) for f in os.listdir(B.OUT_DIR) if f.endswith(".o") and f != "assignment.o"]
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-cudart", "static"])
from icepy4d_b200 import _native
_native.LIB_PATH = lib
import torch
from icepy4d_b200 import ops
sys.path.insert(0, ROOT)
from oracle import sg_oracle
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# correctness on a small ragged problem vs the oracle
g = torch.Generator().manual_seed(3)
Ssm = torch.randn(700, 1023, generator=g) * 3
P = ops.padded_scores(700, 1023, "cuda"); P.copy_(Ssm)
u, v = ops.sinkhorn(P, 1.0, 30)
import math
Pref = sg_oracle.log_optimal_transport(Ssm, torch.tensor(1.0), 30)
Pg = Ssm + u[:-1, None].cpu() + v[None, :-1].cpu() + math.log(700 + 1023)
print("flags", " ".join(flags) or "-", "| max |dP| vs oracle on 700x1023x30:", float((Pg - Pref[:-1, :-1]).abs().max()))
for N in (8192, 8000, 6000):
    S = torch.randn(N, N, device="cuda") * 3
    P = ops.padded_scores(N, N, "cuda"); P.copy_(S)
    ws = ops.AssignWorkspace(N, N, S.device)
    ms = timeit(lambda: ops.sinkhorn(P, 1.0, 100, ws))
    print(f"flags {' '.join(flags) or '-'}: sinkhorn {N}^2 x 100: {ms*10:.2f} us/iter", flush=True)
