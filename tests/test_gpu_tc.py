"""GPU tests of the tensor-core path (tcgen05 GEMM + flash attention) against f64 references computed from the same
bf16-rounded operands, and of the bf16 SuperGlue / LightGlue pipelines against the reference's golden matches."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from icepy4d_b200 import weights  # noqa: E402


@pytest.fixture(scope="module")
def tc():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from icepy4d_b200 import ops_tc
    return ops_tc


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 256), (1000, 768, 256), (257, 512, 512), (100, 1000, 128), (5, 8, 64)])
def test_gemm_tc(tc, M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).bfloat16()
    W = torch.randn(N, K, generator=gen).bfloat16()
    b, R = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
    ref = torch.relu(0.25 * (A.double() @ W.double().t()) + b.double()) + R.double()
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc(A.cuda(), W.cuda(), b.cuda(), residual=R.cuda(), out32=o32, out16=o16, alpha=0.25, relu=True)
    assert torch.allclose(o32.cpu().double(), ref, atol=2e-4 * K ** 0.5, rtol=1e-5)      # f32 accumulation of exact bf16 products
    assert torch.allclose(o16.cpu().double(), ref, atol=0.05, rtol=1e-2)                 # bf16 rounding of the output
    # column-slice operands / outputs (leading dimensions)
    big = torch.zeros(M, K + 64, dtype=torch.bfloat16)
    big[:, :K] = A
    out = torch.zeros(M, N + 16, device="cuda")
    tc.gemm_tc(big.cuda()[:, :K], W.cuda(), None, out32=out[:, :N])
    assert torch.allclose(out[:, :N].cpu().double(), A.double() @ W.double().t(), atol=2e-4 * K ** 0.5, rtol=1e-5)
    assert float(out[:, N:].abs().max()) == 0.0
    # bf16-only output (the merged one-barrier-per-tile epilogue), no bias, into a column slice whose neighbours must stay untouched
    wide = torch.zeros(M, N + 24, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc(A.cuda(), W.cuda(), None, out16=wide[:, :N], alpha=0.25)
    assert torch.allclose(wide[:, :N].cpu().double(), 0.25 * (A.double() @ W.double().t()), atol=0.05, rtol=1e-2)
    assert float(wide[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(4000, 768, 256), (16384, 256, 512), (3333, 520, 128)])
def test_gemm_tc_persistent_many_tiles(tc, M, N, K):
    """More output tiles than SMs: every CTA walks several tiles (double-buffered TMEM accumulators, TMA-store epilogue), with
    the residual read and written IN PLACE (the SuperGlue residual stream) and a bf16 shadow written into a column slice."""
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).bfloat16()
    W = (torch.randn(N, K, generator=gen) / K ** 0.5).bfloat16()
    b, R = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
    ref = A.double() @ W.double().t() + b.double() + R.double()
    x32 = R.cuda().clone()
    wide = torch.zeros(M, N + 256, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc(A.cuda(), W.cuda(), b.cuda(), residual=x32, out32=x32, out16=wide[:, :N])
    assert torch.allclose(x32.cpu().double(), ref, atol=2e-4 * K ** 0.5, rtol=1e-5)
    assert torch.allclose(wide[:, :N].cpu().double(), ref, atol=0.05, rtol=1e-2)
    assert float(wide[:, N:].abs().max()) == 0.0
    # bf16-only output with ReLU, no residual (the QKV / first MLP layer form)
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc(A.cuda(), W.cuda(), b.cuda(), out16=o16, relu=True)
    assert torch.allclose(o16.cpu().double(), torch.relu(A.double() @ W.double().t() + b.double()), atol=0.05, rtol=1e-2)


def _attn_ref(q, k, v, scale=0.125):
    qh, kh, vh = (t.double().view(t.shape[0], 4, 64).permute(1, 0, 2) for t in (q, k, v))
    return (torch.softmax(qh @ kh.transpose(1, 2) * scale, -1) @ vh).permute(1, 0, 2).reshape(q.shape[0], 256)


@pytest.mark.parametrize("n0,n1", [(128, 128), (300, 200), (1000, 777), (64, 1), (513, 1025), (2400, 2300)])   # the last one: 152 work items -> 4 are split along the keys and merged
def test_attention_tc_self_and_cross(tc, n0, n1):
    gen = torch.Generator().manual_seed(n0 * 3 + n1)
    nt = n0 + n1
    X = (torch.randn(nt, 768, generator=gen) * 1.5).bfloat16()
    q, k, v = X[:, :256], X[:, 256:512], X[:, 512:]
    Xd = X.cuda()
    out = torch.zeros(nt, 256, device="cuda", dtype=torch.bfloat16)
    tc.attention_tc(Xd, [(0, n0, 0, n0), (n0, n1, n0, n1)], out, 0, 256, 512)
    ref = torch.cat([_attn_ref(q[:n0], k[:n0], v[:n0]), _attn_ref(q[n0:], k[n0:], v[n0:])])
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 0.03, err                                       # P rounded to bf16 (2^-9 rel.) + bf16 output
    out.zero_()
    tc.attention_tc(Xd, [(0, n0, n0, n1), (n0, n1, 0, n0)], out, 0, 256, 512)
    ref = torch.cat([_attn_ref(q[:n0], k[n0:], v[n0:]), _attn_ref(q[n0:], k[:n0], v[:n0])])
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 0.03, err


def test_attention_tc_peaked_softmax(tc):
    """Large logits (one dominant key per query) exercise the running-max rescaling across key blocks."""
    gen = torch.Generator().manual_seed(9)
    n = 700
    X = torch.randn(n, 768, generator=gen)
    X[:, :256] *= 6.0
    X = X.bfloat16()
    out = torch.zeros(n, 256, device="cuda", dtype=torch.bfloat16)
    tc.attention_tc(X.cuda(), [(0, n, 0, n)], out, 0, 256, 512)
    ref = _attn_ref(X[:, :256], X[:, 256:512], X[:, 512:])
    assert (out.cpu().double() - ref).abs().max().item() < 0.05


def test_superglue_bf16_matches_reference(golden_dir):
    from icepy4d_b200.matching.superglue import SuperGlueB200
    g = np.load(os.path.join(golden_dir, "sp_sg_small.npz"))
    sg = SuperGlueB200(weights.make_superglue_state(2), sinkhorn_iterations=20, match_threshold=0.2, precision="bf16")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    m0, m1, s0, s1 = sg.match(c("kpts0"), c("scores0"), c("desc0").t().contiguous(), g["image0"].shape,
                              c("kpts1"), c("scores1"), c("desc1").t().contiguous(), g["image1"].shape)
    a = {(i, int(j)) for i, j in enumerate(m0.cpu().numpy()) if j >= 0}
    b = {(i, int(j)) for i, j in enumerate(g["matches0"]) if j >= 0}
    iou = len(a & b) / max(1, len(a | b))
    print("SuperGlue bf16 tensor-core match IoU vs reference:", iou, len(a), len(b))
    assert iou >= 0.99, iou                                      # north-star gate


def test_lightglue_bf16_matches_reference(golden_dir):
    from icepy4d_b200.matching.lightglue import LightGlueB200
    g = np.load(os.path.join(golden_dir, "lg_plain.npz"))
    lg = LightGlueB200(weights.make_lightglue_state(3), precision="bf16", depth_confidence=-1, width_confidence=-1)
    c = lambda k: torch.from_numpy(g[k]).cuda()
    out = lg.match(c("kpts0"), c("desc0"), tuple(g["size0"]), c("kpts1"), c("desc1"), tuple(g["size1"]))
    a = {(i, int(j)) for i, j in enumerate(out["matches0"].cpu().numpy()) if j >= 0}
    b = {(i, int(j)) for i, j in enumerate(g["matches0"]) if j >= 0}
    iou = len(a & b) / max(1, len(a | b))
    print("LightGlue bf16 tensor-core match IoU vs reference:", iou, len(a), len(b))
    assert iou >= 0.99, iou


@pytest.mark.parametrize("n0,n1,p0,p1", [(300, 200, 512, 256), (1000, 777, 1024, 1024), (2300, 2050, 2304, 2560), (129, 1, 256, 256)])
def test_attention_tc_device_side_key_counts(tc, n0, n1, p0, p1):
    """Shape buckets: the launch runs at the padded sizes (p0, p1), the real key counts are read on the device.  Real query rows
    must equal the exact-size launch (padding keys hold large garbage, they must not leak; blocks with no real key at all and
    stream-K segments that start in one are covered by the larger cases)."""
    gen = torch.Generator().manual_seed(n0 * 5 + n1)
    X = (torch.randn(p0 + p1, 768, generator=gen) * 1.5).bfloat16()
    X[n0:p0] = 50.0                                                  # padding rows: would dominate every softmax if they leaked
    X[p0 + n1:] = -50.0
    q, k, v = X[:, :256], X[:, 256:512], X[:, 512:]
    Xd = X.cuda()
    out = torch.zeros(p0 + p1, 256, device="cuda", dtype=torch.bfloat16)
    counts = torch.tensor([n0, n1, n1, n0], device="cuda", dtype=torch.int32)
    tc.attention_tc(Xd, [(0, p0, 0, p0), (p0, p1, p0, p1)], out, 0, 256, 512, key_counts=counts[:2])
    ref = torch.cat([_attn_ref(q[:n0], k[:n0], v[:n0]), _attn_ref(q[p0:p0 + n1], k[p0:p0 + n1], v[p0:p0 + n1])])
    got = torch.cat([out[:n0], out[p0:p0 + n1]]).cpu().double()
    assert torch.isfinite(out.float()).all()
    assert (got - ref).abs().max().item() < 0.03
    out.zero_()
    tc.attention_tc(Xd, [(0, p0, p0, p1), (p0, p1, 0, p0)], out, 0, 256, 512, key_counts=counts[2:])
    ref = torch.cat([_attn_ref(q[:n0], k[p0:p0 + n1], v[p0:p0 + n1]), _attn_ref(q[p0:p0 + n1], k[:n0], v[:n0])])
    got = torch.cat([out[:n0], out[p0:p0 + n1]]).cpu().double()
    assert torch.isfinite(out.float()).all()
    assert (got - ref).abs().max().item() < 0.03


def test_superglue_graph_buckets_serve_uneven_keypoint_counts(tc):
    """Tiles of real images have different keypoint counts: the CUDA-graph path is keyed on 256-keypoint buckets and must give
    what the eager schedule gives at the exact sizes (same kernels; only the stream-K split of the attention differs)."""
    from icepy4d_b200 import weights
    from icepy4d_b200.matching.superglue import SuperGlueWeights
    w = SuperGlueWeights(weights.make_superglue_state(2), "cuda")
    net = tc.SuperGlueTensorCore(w, "cuda")
    assert net.use_graphs
    gen = torch.Generator().manual_seed(4)
    sizes = [(1500, 1301), (1410, 1290), (1536, 1280), (1281, 1025)]       # the first three share the (1536, 1536) / (1536, 1280) buckets
    for rep in range(2):
        for n0, n1 in sizes:
            d0 = torch.nn.functional.normalize(torch.randn(n0, 256, generator=gen), dim=1).cuda()
            d1 = torch.nn.functional.normalize(torch.randn(n1, 256, generator=gen), dim=1).cuda()
            got = net.gnn_and_scores(d0, d1).clone()
            type(net).use_graphs = False
            try:
                ref = net.gnn_and_scores(d0, d1).clone()
            finally:
                type(net).use_graphs = True
            assert got.shape == ref.shape == (n0, n1)
            err = (got - ref).abs().max().item()
            agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
            print(f"SuperGlue graph replay vs eager ({n0}, {n1}) rep {rep}: max |dscore| {err:.4f} (|score| max {ref.abs().max().item():.2f}), "
                  f"row arg-max agreement {agree:.4f}")
            assert err < 5e-2 * max(1.0, ref.abs().max().item()), (n0, n1, rep, err)
            assert agree > 0.99, (n0, n1, rep, agree)
    graphs = [k for k, v in net._graphs.items() if isinstance(v, dict)]
    assert len(graphs) >= 2, "bucketed graphs were not captured"
    assert all(k[0] % 256 == 0 and k[1] % 256 == 0 for k in graphs)


def test_lightglue_static_schedule_graph_buckets(tc):
    """The static LightGlue schedule (no early stop, no pruning) is replayed from a CUDA graph keyed on 256-keypoint buckets: for
    uneven keypoint counts the replay must give the eager schedule's matches (same kernels; only the attention's stream-K split
    and the padding rows differ)."""
    from icepy4d_b200 import synthetic
    from icepy4d_b200.matching.lightglue import LightGlueB200
    lg = LightGlueB200(weights.make_lightglue_state(3), precision="bf16", depth_confidence=-1, width_confidence=-1)
    assert lg._tc.use_graphs
    gen = torch.Generator().manual_seed(11)
    for n0, n1 in [(1500, 1301), (1410, 1290), (1536, 1280)]:               # the first two share the (1536, 1536) bucket
        k0 = torch.rand(n0, 2, generator=gen) * torch.tensor([1328.0, 1498.0])
        shift = torch.tensor([16.0, 24.0])
        perm = torch.randperm(n0, generator=gen)[:n1]
        k1 = (k0[perm] + shift) if n1 <= n0 else k0
        d0 = torch.nn.functional.normalize(torch.randn(n0, 256, generator=gen), dim=1)
        d1 = torch.nn.functional.normalize(d0[perm] + 0.15 * torch.randn(n1, 256, generator=gen), dim=1)
        args = (k0.cuda(), d0.cuda(), (1329, 1499), k1.cuda(), d1.cuda(), (1329, 1499))
        runs = [lg.match(*args) for _ in range(3)]                            # eager (bucket seen), capture + replay, replay
        type(lg._tc).use_graphs = False
        try:
            ref = lg.match(*args)
        finally:
            type(lg._tc).use_graphs = True
        r0 = ref["matches0"].cpu()
        assert (r0 > -1).sum() > 0.5 * n1
        for out in runs:
            m0 = out["matches0"].cpu()
            assert m0.shape == r0.shape
            agree = (m0 == r0).float().mean().item()
            print(f"LightGlue graph replay vs eager ({n0}, {n1}): matches0 agreement {agree:.4f}")
            assert agree > 0.99, (n0, n1, agree)
            same = m0 == r0                                                   # a flipped borderline match changes its score to / from 0
            ds = (out["matching_scores0"].cpu()[same] - ref["matching_scores0"].cpu()[same]).abs().max().item()
            assert ds < 0.1, (n0, n1, ds)
    graphs = [k for k, v in lg._tc._graphs.items() if isinstance(v, dict)]
    assert len(graphs) >= 2 and all(k[0] % 256 == 0 and k[1] % 256 == 0 for k in graphs), "bucketed graphs were not captured"


@pytest.mark.parametrize("n", [128, 1000, 4133])
def test_gemm_tc_rotary_epilogue_equals_gemm_then_rotary_pass(tc, n):
    """LightGlue's fused QKV projection: the rotary embedding of q and k in the GEMM epilogue must give what the two-kernel path
    (GEMM -> f32 -> i4d_lg_rotary_cast_bf16) gives — same f32 arithmetic, one rounding to bf16."""
    gen = torch.Generator().manual_seed(n)
    A = (torch.randn(n, 256, generator=gen)).bfloat16().cuda()
    W = (torch.randn(768, 256, generator=gen) * 0.1).bfloat16().cuda()
    b = torch.randn(768, generator=gen).cuda()
    ang = torch.rand(n, 32, generator=gen) * 6.28
    cs = torch.cat([torch.cos(ang), torch.sin(ang)], 1).contiguous().cuda()
    q32 = torch.empty(n, 768, device="cuda")
    ref = torch.empty(n, 768, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc(A, W, b, out32=q32)
    tc.rotary_cast_bf16(q32, cs, ref)
    out = torch.empty(n, 768, device="cuda", dtype=torch.bfloat16)
    tc.gemm_tc_rotary(A, W, b, cs, 512, out)
    d = (out.float() - ref.float()).abs()
    assert torch.equal(out[:, 512:], ref[:, 512:])                           # v: untouched
    assert d.max().item() <= 2.0 ** -7 * ref.float().abs().max().item()      # at most one bf16 ulp (fma contraction may differ)
    assert (out != ref).float().mean().item() < 1e-2


def test_attention_tc_is_bit_reproducible(tc):
    """Items cut by the stream-K range boundaries are merged by whichever part arrives last; the parts are folded in a fixed order,
    so repeated launches give identical bits (2400 / 2300 keypoints: several split items)."""
    gen = torch.Generator().manual_seed(77)
    n0, n1 = 2400, 2300
    X = (torch.randn(n0 + n1, 768, generator=gen) * 1.5).bfloat16().cuda()
    outs = []
    for _ in range(6):
        out = torch.zeros(n0 + n1, 256, device="cuda", dtype=torch.bfloat16)
        tc.attention_tc(X, [(0, n0, n0, n1), (n0, n1, 0, n0)], out, 0, 256, 512)
        outs.append(out)
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])


@pytest.mark.parametrize("n", [1, 77, 2048, 8192])
def test_superglue_keypoint_encoder_split_precision_gemms(tc, n):
    """The wide layers of SuperGlue's KeypointEncoder (superglue.py:51-61,67-78) as three-product split-precision tensor-core GEMMs
    ([hi|hi|lo] x [hi|lo|hi] through the bf16 GEMM, f32 accumulation) against the f32 SIMT path and an f64 evaluation of the same
    folded weights: the split path must be as close to f64 as the f32 path is (both ~1e-6 relative), i.e. invisible next to the bf16
    rounding (4e-3) the GNN applies to the encoder's output."""
    from icepy4d_b200 import weights
    from icepy4d_b200.matching.superglue import SuperGlueB200
    sg = SuperGlueB200(weights.make_superglue_state(2), precision="bf16", sinkhorn_iterations=5)
    gen = torch.Generator().manual_seed(n)
    kpts = (torch.rand(n, 2, generator=gen) * torch.tensor([1999.0, 1332.0])).cuda()
    sc = torch.rand(n, generator=gen).cuda()
    desc = torch.nn.functional.normalize(torch.randn(n, 256, generator=gen), dim=1).cuda()
    ref32 = sg.encode(kpts, sc, desc, 1333, 2000)                       # five f32 SIMT GEMMs
    out = sg._tc.encode(kpts, sc, desc, 1333, 2000)
    from icepy4d_b200 import ops
    x = ops.sg_kenc_input(kpts, sc, 2000.0, 1333.0).double()
    for i, (w, b) in enumerate(sg.w.kenc):
        x = x @ w.double().t() + b.double()
        if i < 4:
            x = torch.relu(x)
    ref64 = x + desc.double()
    scale = ref64.abs().max().item()
    e32 = (ref32.double() - ref64).abs().max().item() / scale
    e3 = (out.double() - ref64).abs().max().item() / scale
    assert e32 < 2e-5, e32
    assert e3 < 2e-5, (e3, e32)
    # the weight / activation layouts: [hi | lo | hi] and [hi | hi | lo], lo = bf16(x - hi)
    w = torch.randn(64, 128, generator=gen).cuda()
    for weight in (False, True):
        s3 = tc.split3_bf16(w, torch.empty(64, 384, device="cuda", dtype=torch.bfloat16), weight=weight)
        hi = w.bfloat16()
        lo = (w - hi.float()).bfloat16()
        exp = torch.cat([hi, lo, hi] if weight else [hi, hi, lo], 1)
        assert torch.equal(s3, exp)
