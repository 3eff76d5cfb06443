"""Per-kernel timings at cfg2 sizes with CUDA events (not under a profiler).  Prints one line per kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops, ops_tc, weights, synthetic
from icepy4d_b200.matching.superpoint import SuperPointB200

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {}
N = 8192
nt = 2 * N
dev = "cuda"
# attention: both images, 4 heads
qkv = (torch.randn(nt, 768, device=dev)).bfloat16()
att = torch.empty(nt, 256, device=dev, dtype=torch.bfloat16)
probs = [(0, N, 0, N), (N, N, N, N)]
ms = timeit(lambda: ops_tc.attention_tc(qkv, probs, att, 0, 256, 512))
fl = 2 * 4 * (2 * N * N * 64 * 2)
res["attn_tc (2 img x 4 heads, 8192^2)"] = (ms, f"{fl/ms/1e9:.1f} TFLOP/s")
# gemms
x = torch.randn(nt, 512, device=dev).bfloat16()
for (K, Nn) in ((256, 768), (256, 256), (512, 512), (512, 256)):
    W = torch.randn(Nn, K, device=dev).bfloat16(); b = torch.randn(Nn, device=dev)
    o = torch.empty(nt, Nn, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops_tc.gemm_tc(x[:, :K], W, b, out16=o))
    res[f"gemm_tc 16384x{Nn}x{K}"] = (ms, f"{2*nt*Nn*K/ms/1e9:.1f} TFLOP/s")
md = torch.randn(nt, 256, device=dev).bfloat16(); sc = torch.empty(N, N, device=dev)
ms = timeit(lambda: ops_tc.gemm_tc(md[:N], md[N:], out32=sc, alpha=1/16))
res["gemm_tc scores 8192x8192x256 (f32 out)"] = (ms, f"{2*N*N*256/ms/1e9:.1f} TFLOP/s, {N*N*4/ms/1e6:.0f} GB/s write")
# sinkhorn pieces
S = torch.randn(N, N, device=dev); ws = ops.AssignWorkspace(N, N, S.device)
v = torch.zeros(N, device=dev); u = torch.zeros(N, device=dev)
ms = timeit(lambda: ops.row_lse(S, 1.0, v)); res["row_lse 8192^2"] = (ms, f"{N*N*4/ms/1e6:.0f} GB/s")
ms = timeit(lambda: ops.col_lse(S, 1.0, u, ws)); res["col_lse 8192^2 (partial+combine)"] = (ms, f"{N*N*4/ms/1e6:.0f} GB/s")
ms = timeit(lambda: ops.sinkhorn(S, 1.0, 10, ws), reps=5); res["sinkhorn 10 iters 8192^2"] = (ms, f"{ms/10*1000:.1f} us/iter, {2*N*N*4*10/ms/1e6:.0f} GB/s algorithmic")
ms = timeit(lambda: ops.sg_assign(S, 1.0, 100, 0.2, ws), reps=3, warm=1); res["sg_assign 100 iters 8192^2"] = (ms, f"{ms/100*1000:.1f} us/iter")
ms = timeit(lambda: ops.sg_assign(S, 1.0, 0, 0.2, ws)); res["sg_assign 0 iters 8192^2 (row+col arg-max in one read, mutual NN)"] = (ms, f"{N*N*4/ms/1e6:.0f} GB/s on one read")
z0 = torch.randn(N, device=dev); z1 = torch.randn(N, device=dev)
ms = timeit(lambda: ops.lg_assign(S, z0, z1, 0.1, ws)); res["lg_assign 8192^2 (double softmax + mutual NN, two reads of sim)"] = (ms, f"{2*N*N*4/ms/1e6:.0f} GB/s on two reads")
if os.environ.get("KB_LG16K", "1") == "1":
    N2 = 16384
    S2 = torch.randn(N2, N2, device=dev); ws2 = ops.AssignWorkspace(N2, N2, S2.device)
    z0 = torch.randn(N2, device=dev); z1 = torch.randn(N2, device=dev)
    ms = timeit(lambda: ops.lg_assign(S2, z0, z1, 0.1, ws2), reps=5); res["lg_assign 16384^2 (two reads of sim)"] = (ms, f"{2*N2*N2*4/ms/1e6:.0f} GB/s on two reads")
    del S2, ws2
# superpoint
i0, _ = synthetic.stereo_pair(1999, 1999, seed=1, channels=1)
img = torch.tensor(i0 / 255.0, dtype=torch.float)[None, None].cuda()
for prec in ("tf32", "bf16"):
    sp = SuperPointB200(weights.make_superpoint_state(1), nms_radius=3, keypoint_threshold=1e-4, max_keypoints=8192, conv_precision=prec)
    ms = timeit(lambda: sp.backbone(img), reps=5); res[f"superpoint backbone 1999^2 ({prec})"] = (ms, f"{676.4/ms:.1f} TFLOP/s")
    logits, desc = sp.backbone(img)
ms = timeit(lambda: ops.sp_score_map(logits)); res["sp_score_map"] = (ms, f"{(65+64)*249*249*4/ms/1e6:.0f} GB/s")
scores = ops.sp_score_map(logits)
kws = ops.KeypointWorkspace(scores.shape[0], scores.shape[1], scores.device)
ms = timeit(lambda: ops.sp_keypoints(scores, 3, 1e-4, 4, 8192, kws)); res["sp nms+topk 1992^2"] = (ms, "")
from icepy4d_b200 import _native as NN
st = NN.current_stream()
ms = timeit(lambda: NN.call("i4d_sp_nms_candidates", scores, scores.shape[0], scores.shape[1], 3, 1e-4, 4, kws.keys, kws.cand_cap, kws.count, None, st)); res["  nms only"] = (ms, "")
_k = torch.empty((8192, 2), device="cuda"); _s = torch.empty(8192, device="cuda"); _n = torch.zeros(1, device="cuda", dtype=torch.int32)
ms = timeit(lambda: NN.call("i4d_sp_select_topk", kws.keys, kws.count, kws.cand_cap, 8192, scores.shape[1], _k, _s, 8192, _n, kws.spill, st)); res["  topk only"] = (ms, "")
kp, ks, n, _ = ops.sp_keypoints(scores, 3, 1e-4, 4, 8192, kws)
print("candidates", int(kws.count.item()), "kept", int(n.item()))
ms = timeit(lambda: ops.sp_sample_descriptors(desc, kp, n)); res["sp_sample_descriptors 8192"] = (ms, f"{8192*5*1024/ms/1e6:.0f} GB/s")
ms = timeit(lambda: sp.postprocess(logits, desc)); res["sp postprocess total"] = (ms, "")
for k, (ms, note) in res.items():
    print(f"{ms*1000:10.1f} us  {k:50s} {note}")
