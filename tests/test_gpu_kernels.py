"""GPU parity tests: every CUDA kernel, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact for index/compare work (NMS, top-k, argmax, status codes); tolerances stated per test for floating point."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from icepy4d_b200 import synthetic, weights  # noqa: E402
from oracle import geom_oracle, lg_oracle, sg_oracle, sp_oracle  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from icepy4d_b200 import ops as _ops
    return _ops


def _set(k):
    return {(float(x), float(y)) for x, y in k}


# ------------------------------------------------------------------ SuperPoint post-processing
def test_score_map(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "sp_sg_small.npz"))
    logits = torch.from_numpy(g["logits0"])
    ref = sp_oracle.score_map(logits)
    out = ops.sp_score_map(logits.cuda()).cpu()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-8)   # expf vs CPU vectorised exp: a few ulp


def test_nms_teacher_forced_identical_keypoints(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "sp_sg_small.npz"))
    scores = sp_oracle.score_map(torch.from_numpy(g["logits0"]))          # oracle's own f32 score map (this host)
    ref_nms = sp_oracle.simple_nms(scores, 3)
    rk, rs = sp_oracle.keypoints_sg(ref_nms, 1e-4, 4, 256)
    kp, sc, n, nms = ops.sp_keypoints(scores.cuda(), 3, 1e-4, 4, 256, want_nms=True)
    assert torch.equal(nms.cpu(), ref_nms)                                 # whole NMS map bit-identical
    n = int(n.item())
    assert n == len(rk)
    assert _set(kp[:n].cpu().numpy()) == _set(rk.numpy())                  # identical keypoint indices
    ref = {(float(x), float(y)): float(s) for (x, y), s in zip(rk.numpy(), rs.numpy())}
    for (x, y), s in zip(kp[:n].cpu().numpy(), sc[:n].cpu().numpy()):
        assert ref[(float(x), float(y))] == float(s)
    assert np.all(np.diff(sc[:n].cpu().numpy()) <= 0)                      # score-descending like torch.topk
    # against the reference's own output (generated on another host: CPU softmax may differ in the last ulp,
    # which can only swap near-tied candidates at the top-k boundary)
    a, b = _set(kp[:n].cpu().numpy()), _set(g["kpts0"])
    assert len(a & b) / len(a | b) >= 0.99


@pytest.mark.parametrize("H,W,r", [(64, 64, 1), (200, 333, 2), (129, 257, 3), (300, 190, 4), (8, 8, 4), (70, 500, 0)])
def test_nms_random_with_ties(ops, H, W, r):
    gen = torch.Generator().manual_seed(H * 1000 + W + r)
    s = torch.randint(0, 12, (H, W), generator=gen).float() / 12.0          # heavy ties / plateaus
    s[torch.rand(H, W, generator=gen) < 0.3] = 0.0
    ref = sp_oracle.simple_nms(s, r)
    kp, sc, n, nms = ops.sp_keypoints(s.cuda(), r, 0.05, 4, -1, want_nms=True)
    assert torch.equal(nms.cpu(), ref)
    rk, rs = sp_oracle.keypoints_sg(ref, 0.05, 4, -1)
    n = int(n.item())
    assert n == len(rk)
    assert torch.equal(kp[:n].cpu(), rk) and torch.equal(sc[:n].cpu(), rs)  # keep-all => row-major order, exact


def test_topk_edge_cases(ops):
    gen = torch.Generator().manual_seed(5)
    s = torch.rand(96, 160, generator=gen)
    ref = sp_oracle.simple_nms(s, 2)
    rk, rs = sp_oracle.keypoints_sg(ref, 0.5, 4, -1)
    for k in (1, 7, len(rk) - 1, len(rk), len(rk) + 5, 4096):
        kp, sc, n, _ = ops.sp_keypoints(s.cuda(), 2, 0.5, 4, k)
        n = int(n.item())
        assert n == min(k, len(rk))
        ek, es = sp_oracle.keypoints_sg(ref, 0.5, 4, k)
        assert _set(kp[:n].cpu().numpy()) == _set(ek.numpy())
        assert np.array_equal(np.sort(sc[:n].cpu().numpy()), np.sort(es.numpy()))
    # empty result
    kp, sc, n, _ = ops.sp_keypoints(torch.zeros(64, 64).cuda(), 3, 0.5, 4, 100)
    assert int(n.item()) == 0


def test_topk_large_keep_all(ops):
    """> 16384 survivors exercises the global-memory sort path."""
    gen = torch.Generator().manual_seed(6)
    s = torch.rand(600, 700, generator=gen)
    ref = sp_oracle.simple_nms(s, 1)
    rk, rs = sp_oracle.keypoints_sg(ref, 0.1, 4, -1)
    assert len(rk) > 16384
    kp, sc, n, _ = ops.sp_keypoints(s.cuda(), 1, 0.1, 4, -1)
    n = int(n.item())
    assert n == len(rk) and torch.equal(kp[:n].cpu(), rk) and torch.equal(sc[:n].cpu(), rs)
    k = 20000
    kp, sc, n, _ = ops.sp_keypoints(s.cuda(), 1, 0.1, 4, k)
    ek, es = sp_oracle.keypoints_sg(ref, 0.1, 4, k)
    assert int(n.item()) == k and _set(kp[:k].cpu().numpy()) == _set(ek.numpy())


def test_sample_descriptors(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "sp_sg_small.npz"))
    sd = weights.make_superpoint_state(1)
    _, desc = sp_oracle.backbone(torch.tensor(g["image0"] / 255.0, dtype=torch.float)[None, None], sd)
    kp = torch.from_numpy(g["kpts0"])
    # add border / out-of-range taps and fractional positions
    extra = torch.tensor([[0.0, 0.0], [319.0, 239.0], [3.5, 3.5], [100.25, 57.75], [318.9, 0.1]])
    kp = torch.cat([kp, extra])
    ref = sp_oracle.sample_descriptors(kp, desc).t()
    out = ops.sp_sample_descriptors(desc.permute(1, 2, 0).contiguous().cuda(), kp.cuda()).cpu()
    assert torch.allclose(out, ref, atol=2e-6)
    assert np.allclose(out[: len(g["kpts0"])].numpy().T, g["desc0"], atol=2e-5)   # vs the reference itself


# ------------------------------------------------------------------ dense f32
@pytest.mark.parametrize("M,N,K", [(1, 1, 3), (100, 32, 3), (257, 65, 19), (512, 768, 256), (300, 256, 512)])
def test_gemm_f32(ops, M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen)
    b, R = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
    ref = torch.relu(0.5 * (A.double() @ W.double().t()) + b.double()) + R.double()
    out = ops.gemm_f32(A.cuda(), W.cuda(), b.cuda(), residual=R.cuda(), alpha=0.5, relu=True).cpu()
    assert torch.allclose(out.double(), ref, atol=1e-4 * K ** 0.5)


@pytest.mark.parametrize("Nq,Nk", [(64, 64), (100, 333), (257, 1), (300, 129)])
def test_attention_f32(ops, Nq, Nk):
    gen = torch.Generator().manual_seed(Nq * 7 + Nk)
    q, k, v = (torch.randn(n, 256, generator=gen) for n in (Nq, Nk, Nk))
    qh, kh, vh = (t.view(-1, 4, 64).permute(1, 0, 2).double() for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(1, 2) / 8, -1) @ vh).permute(1, 0, 2).reshape(Nq, 256)
    out = torch.empty(Nq, 256, device="cuda")
    ops.attention_f32(q.cuda(), k.cuda(), v.cuda(), out)
    assert torch.allclose(out.cpu().double(), ref, atol=2e-5)


def test_layernorm_gelu_posenc_rotary(ops):
    gen = torch.Generator().manual_seed(11)
    x, g_, b_ = torch.randn(77, 512, generator=gen), torch.rand(512, generator=gen) + 0.5, torch.randn(512, generator=gen)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x, (512,), g_, b_))
    out = ops.layernorm_gelu(x.cuda().clone(), g_.cuda(), b_.cuda()).cpu()
    assert torch.allclose(out, ref, atol=2e-5)
    sd = weights.make_lightglue_state(3)
    kp = torch.rand(50, 2, generator=gen) * torch.tensor([640.0, 480.0])
    cs_ref = lg_oracle.posenc(lg_oracle.normalize_keypoints(kp, (640.0, 480.0)), sd)
    cs = ops.lg_posenc(kp.cuda(), 640.0, 480.0, sd["posenc.Wr.weight"].cuda())
    assert torch.allclose(cs[:, :32].cpu(), cs_ref[0][:, ::2], atol=1e-5)
    assert torch.allclose(cs[:, 32:].cpu(), cs_ref[1][:, ::2], atol=1e-5)
    t = torch.randn(50, 256, generator=gen)
    ref = lg_oracle.rotary(t.view(50, 4, 64).permute(1, 0, 2), cs_ref).permute(1, 0, 2).reshape(50, 256)
    out = ops.lg_rotary_(t.cuda().clone(), cs).cpu()
    assert torch.allclose(out, ref, atol=1e-5)


# ------------------------------------------------------------------ assignment
@pytest.mark.parametrize("M,N", [(1, 1), (37, 53), (256, 256), (300, 1001), (1024, 516)])
def test_row_col_lse(ops, M, N):
    gen = torch.Generator().manual_seed(M * 3 + N)
    S = torch.randn(M, N, generator=gen) * 5
    co, ro = torch.randn(N, generator=gen), torch.randn(M, generator=gen)
    r = ops.row_lse(S.cuda(), 2.0, co.cuda()).cpu()
    c = ops.col_lse(S.cuda(), 1.0, ro.cuda()).cpu()
    assert torch.allclose(r, torch.logsumexp(2 * S + co[None], 1), atol=1e-4)
    assert torch.allclose(c, torch.logsumexp(S + ro[:, None], 0), atol=1e-4)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,iters", [(5, 7, 3), (200, 256, 20), (513, 300, 100), (1, 64, 4), (2500, 4096, 30), (700, 8192, 5)])
def test_sinkhorn_and_sg_assign(ops, M, N, iters, mode):
    """mode 0: fused persistent kernel, a-priori stabilisers after 2 iterations; mode 1: two-pass kernels; mode 2: fused, running maxima."""
    ops.set_sinkhorn_mode(mode)
    try:
        _sinkhorn_case(ops, M, N, iters)
    finally:
        ops.set_sinkhorn_mode(0)


def _sinkhorn_case(ops, M, N, iters):
    gen = torch.Generator().manual_seed(M + N)
    S = torch.randn(M, N, generator=gen) * 3
    for i in range(0, min(M, N), 2):
        S[i, (i * 7) % N] += 25.0                               # planted matches
    alpha = torch.tensor(1.0)
    P = sg_oracle.log_optimal_transport(S, alpha, iters)
    u, v = ops.sinkhorn(S.cuda(), 1.0, iters)
    norm = -np.log(M + N)
    Z = torch.full((M + 1, N + 1), 1.0)
    Z[:M, :N] = S
    Pg = Z + u.cpu()[:, None] + v.cpu()[None, :] - norm
    assert torch.allclose(Pg, P, atol=2e-3)                      # f32 LSE in a different summation order, 100 iterations
    m0, m1, s0, s1 = (t.cpu() for t in ops.sg_assign(S.cuda(), 1.0, iters, 0.2))
    r0, r1, rs0, rs1 = sg_oracle.mutual_nn(P, 0.2)
    assert torch.equal(m0.long(), r0) and torch.equal(m1.long(), r1)
    assert torch.allclose(s0, rs0, atol=1e-3) and torch.allclose(s1, rs1, atol=1e-3)


@pytest.mark.parametrize("M,N,iters", [(8192, 8192, 12), (8191, 8188, 7), (4100, 8192, 9), (3000, 2048, 11)])
def test_sinkhorn_fused_vs_two_pass_large(ops, M, N, iters):
    """Full cfg2 size (and ragged bands / partial column groups): the fused persistent kernel (boustrophedon bands,
    mbarrier-decoupled warps, a-priori stabilisers) against the plain two-pass row/column kernels on the same matrix."""
    gen = torch.Generator().manual_seed(M * 7 + N)
    S = (torch.randn(M, N, generator=gen) * 3).cuda()
    idx = torch.arange(0, min(M, N), 2)
    S[idx, (idx * 7) % N] += 25.0
    ops.set_sinkhorn_mode(1)
    try:
        u1, v1 = ops.sinkhorn(S, 1.0, iters)
        m1 = [t.clone() for t in ops.sg_assign(S, 1.0, iters, 0.2)]
    finally:
        ops.set_sinkhorn_mode(0)
    for mode in (0, 2):
        ops.set_sinkhorn_mode(mode)
        try:
            u0, v0 = ops.sinkhorn(S, 1.0, iters)
            m0 = ops.sg_assign(S, 1.0, iters, 0.2)
        finally:
            ops.set_sinkhorn_mode(0)
        assert torch.isfinite(u0).all() and torch.isfinite(v0).all()
        assert float((u0 - u1).abs().max()) < 2e-3 and float((v0 - v1).abs().max()) < 2e-3, mode
        assert torch.equal(m0[0], m1[0]) and torch.equal(m0[1], m1[1]), mode
        assert torch.allclose(m0[2], m1[2], atol=1e-3)


def _pitched(ops, S):
    P = ops.padded_scores(S.shape[0], S.shape[1], "cuda")
    P.copy_(S)
    return P


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("M,N,iters", [(301, 301, 25), (513, 1023, 30), (700, 8189, 6), (2047, 4097, 12), (64, 66, 5)])
def test_sinkhorn_ragged_n_takes_the_fused_kernel(ops, M, N, iters, mode):
    """Real tiles have arbitrary keypoint counts.  With a pitched score matrix (row pitch = N rounded up to 4 floats, what
    the matchers allocate) N % 4 != 0 runs the fused kernel (pad columns set to -1e30 by the launcher): same potentials and
    matches as the CPU oracle (superglue.py:152-186, :288-298), and exactly one cooperative launch + the pad fill."""
    from icepy4d_b200 import _native
    gen = torch.Generator().manual_seed(M * 11 + N)
    S = torch.randn(M, N, generator=gen) * 3
    for i in range(0, min(M, N), 2):
        S[i, (i * 7) % N] += 25.0
    P = sg_oracle.log_optimal_transport(S, torch.tensor(1.0), iters)
    r0, r1, rs0, rs1 = sg_oracle.mutual_nn(P, 0.2)
    ops.set_sinkhorn_mode(mode)
    try:
        Sp = _pitched(ops, S)
        assert Sp.stride(0) % 4 == 0 and Sp.stride(0) >= N
        n0 = _native.LAUNCHES
        u, v = ops.sinkhorn(Sp, 1.0, iters)
        assert _native.LAUNCHES - n0 == (2 if N % 4 else 1)
        assert u.shape == (M + 1,) and v.shape == (N + 1,)
        m0, m1, s0, s1 = (t.cpu() for t in ops.sg_assign(_pitched(ops, S), 1.0, iters, 0.2))
    finally:
        ops.set_sinkhorn_mode(0)
    Z = torch.full((M + 1, N + 1), 1.0)
    Z[:M, :N] = S
    Pg = Z + u.cpu()[:, None] + v.cpu()[None, :] + np.log(M + N)
    assert torch.allclose(Pg, P, atol=2e-3)
    assert torch.equal(m0.long(), r0) and torch.equal(m1.long(), r1)
    assert torch.allclose(s0, rs0, atol=1e-3) and torch.allclose(s1, rs1, atol=1e-3)


def test_sinkhorn_fast_mode_restart(ops):
    """Potentials that jump by hundreds of nats between iterations underflow the a-priori stabilisers: the fused kernel
    must notice, restart in exact mode on the device and still agree with the two-pass kernels."""
    M, N, iters = 600, 1024, 8
    gen = torch.Generator().manual_seed(5)
    S = (torch.randn(M, N, generator=gen) * 60).cuda()
    S[::3] -= 400.0
    S[:, ::5] += 300.0
    ops.set_sinkhorn_mode(1)
    try:
        u1, v1 = ops.sinkhorn(S, 1.0, iters)
    finally:
        ops.set_sinkhorn_mode(0)
    u0, v0 = ops.sinkhorn(S, 1.0, iters)
    assert torch.isfinite(u0).all() and torch.isfinite(v0).all()
    assert float((u0 - u1).abs().max()) < 5e-2 and float((v0 - v1).abs().max()) < 5e-2


@pytest.mark.parametrize("M,N", [(3, 5), (256, 200), (700, 513), (1030, 2051)])
def test_lg_assign(ops, M, N):
    gen = torch.Generator().manual_seed(M * 5 + N)
    sim = torch.randn(M, N, generator=gen) * 2
    for i in range(0, min(M, N), 2):
        sim[i, (i * 3) % N] += 20.0
    z0, z1 = torch.randn(M, 1, generator=gen) * 3, torch.randn(N, 1, generator=gen) * 3
    F = torch.nn.functional
    P = sim.new_zeros(M + 1, N + 1)
    P[:M, :N] = F.log_softmax(sim, 1) + F.log_softmax(sim, 0) + F.logsigmoid(z0) + F.logsigmoid(z1).t()
    r0, r1, rs0, rs1 = sg_oracle.mutual_nn(P, 0.1)
    for dev_sim in (sim.cuda(), _pitched(ops, sim)):           # contiguous (scalar tail) and pitched (128-bit loads) layouts
        m0, m1, s0, s1 = (t.cpu() for t in ops.lg_assign(dev_sim, z0.cuda(), z1.cuda(), 0.1))
        assert torch.equal(m0.long(), r0) and torch.equal(m1.long(), r1)
        assert torch.allclose(s0, rs0, atol=1e-4) and torch.allclose(s1, rs1, atol=1e-4)


@pytest.mark.parametrize("M,N,spread", [(700, 513, False), (1030, 2051, True), (129, 640, True)])
def test_lg_assign_one_read_passes_and_their_exact_fallback(ops, M, N, spread):
    """The double softmax reads `sim` twice in total (row + column LSE in one pass, row + column arg-max in another).  With
    `spread`, one column and one row sit 200 nats below everything else: the shared-exponential column sums of the LSE pass
    underflow there, the device flag goes up and the exact two-pass kernels redo the LSE — the result must still be torch's."""
    gen = torch.Generator().manual_seed(M + 7 * N)
    sim = torch.randn(M, N, generator=gen) * 3
    for i in range(0, min(M, N), 3):
        sim[i, (i * 5) % N] += 15.0
    if spread:
        sim[:, N // 3] -= 200.0
        sim[M // 2, :] -= 200.0
    z0, z1 = torch.randn(M, 1, generator=gen) * 3, torch.randn(N, 1, generator=gen) * 3
    F = torch.nn.functional
    P = sim.new_zeros(M + 1, N + 1)
    P[:M, :N] = (F.log_softmax(sim.double(), 1) + F.log_softmax(sim.double(), 0) + F.logsigmoid(z0.double()) + F.logsigmoid(z1.double()).t()).float()
    r0, r1, rs0, rs1 = sg_oracle.mutual_nn(P, 0.1)
    m0, m1, s0, s1 = (t.cpu() for t in ops.lg_assign(_pitched(ops, sim), z0.cuda(), z1.cuda(), 0.1))
    assert torch.equal(m0.long(), r0) and torch.equal(m1.long(), r1)
    assert torch.allclose(s0, rs0, atol=1e-4) and torch.allclose(s1, rs1, atol=1e-4)


# ------------------------------------------------------------------ geometry
def test_undistort_bit_exact_vs_opencv(ops):
    sc = synthetic.two_view_scene(n=20000, seed=3)
    for pts, K, d in ((sc["pts0"], synthetic.CAM1_K, synthetic.CAM1_DIST), (sc["pts1"], synthetic.CAM2_K, synthetic.CAM2_DIST)):
        ref = geom_oracle.undistort_points(pts, K, d)
        out = ops.undistort_points(torch.from_numpy(pts).cuda(), K, d).cpu().numpy()
        assert np.array_equal(out, ref)


def test_triangulation_vs_golden_and_oracle(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    u0, u1 = torch.from_numpy(g["und0"]).cuda(), torch.from_numpy(g["und1"]).cuda()
    X, st = ops.triangulate_iterative_ls(u0, u1, g["P0"], g["P1"])
    rel = np.linalg.norm(X.cpu().numpy() - g["X_iter"], axis=1) / np.linalg.norm(g["X_iter"], axis=1)
    assert rel.max() < 1e-9                                       # north-star tolerance is 1e-4
    assert np.array_equal(st.cpu().numpy(), g["status"])
    Xl = ops.triangulate_dlt(u0, u1, g["P0"], g["P1"]).cpu().numpy()
    rel = np.linalg.norm(Xl - g["X_lin"], axis=1) / np.linalg.norm(g["X_lin"], axis=1)
    assert rel.max() < 1e-8
    # points behind a camera -> status codes
    P1 = g["P1"].copy()
    P1[2] *= -1
    Xo, so = geom_oracle.iterative_ls(g["und0"][:50], g["P0"], g["und1"][:50], P1)
    Xg, sg_ = ops.triangulate_iterative_ls(u0[:50].contiguous(), u1[:50].contiguous(), g["P0"], P1)
    assert np.array_equal(sg_.cpu().numpy(), so)


# ------------------------------------------------------------------ SuperGlue end to end (f32 path)
def test_superglue_f32_matches_reference(golden_dir):
    from icepy4d_b200.matching.superglue import SuperGlueB200
    g = np.load(os.path.join(golden_dir, "sp_sg_small.npz"))
    sg = SuperGlueB200(weights.make_superglue_state(2), sinkhorn_iterations=20, match_threshold=0.2, precision="f32")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    m0, m1, s0, s1 = sg.match(c("kpts0"), c("scores0"), c("desc0").t().contiguous(), g["image0"].shape,
                              c("kpts1"), c("scores1"), c("desc1").t().contiguous(), g["image1"].shape)
    assert np.array_equal(m0.cpu().numpy(), g["matches0"])        # identical match set vs the reference
    assert np.array_equal(m1.cpu().numpy(), g["matches1"])
    assert np.allclose(s0.cpu().numpy(), g["mscores0"], atol=1e-3)


# ------------------------------------------------------------------ SuperPoint first layer / pooling
@pytest.mark.parametrize("H,W", [(37, 53), (240, 320), (129, 64)])
def test_conv1a_relu_and_maxpool(ops, H, W):
    sd = weights.make_superpoint_state(1)
    gen = torch.Generator().manual_seed(H + W)
    img = torch.rand(1, 1, H, W, generator=gen)
    ref = torch.relu(torch.nn.functional.conv2d(img, sd["conv1a.weight"], sd["conv1a.bias"] + 0.01, padding=1))
    w, b = sd["conv1a.weight"].cuda().contiguous(), (sd["conv1a.bias"] + 0.01).cuda()
    out = ops.sp_conv1a_relu(img.cuda(), w, b, torch.float32)
    assert out.shape == (1, 64, H, W) and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(out.cpu(), ref, atol=2e-6)                          # 9 f32 FMAs per output
    o16 = ops.sp_conv1a_relu(img.cuda(), w, b, torch.float16)
    assert torch.allclose(o16.float().cpu(), ref, atol=2e-3)
    # pooling is exact (pure comparisons), f32 and 16-bit, odd sizes floor like torch
    p = ops.maxpool2x2_cl(out)
    assert torch.equal(p.cpu(), torch.nn.functional.max_pool2d(out.cpu(), 2, 2))
    p16 = ops.maxpool2x2_cl(o16)
    assert torch.equal(p16.cpu(), torch.nn.functional.max_pool2d(o16.cpu().float(), 2, 2).half())
    ob = ops.sp_conv1a_relu(img.cuda(), w, b, torch.bfloat16)
    assert torch.equal(ops.maxpool2x2_cl(ob).cpu(), torch.nn.functional.max_pool2d(ob.cpu().float(), 2, 2).bfloat16())
