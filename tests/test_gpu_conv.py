"""GPU tests of the split-bf16 ("bf16x3") tcgen05 convolutions against f64 convolutions of the same f32 inputs, and of the
whole SuperPoint backbone in that mode against the f32 cuDNN backbone."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from icepy4d_b200 import weights  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from icepy4d_b200 import ops
    return ops


def _split(x, dtype=torch.bfloat16):
    hi = x.to(dtype)
    return torch.stack([hi, (x - hi.float()).to(dtype)])


def _join(p):
    return p[0].float() + p[1].float()


@pytest.mark.parametrize("H,W,cin,cout,k,pool", [
    (16, 16, 64, 64, 3, False), (37, 53, 64, 64, 3, False), (38, 54, 64, 64, 3, True), (37, 53, 64, 64, 3, True),
    (33, 20, 64, 128, 3, False), (21, 40, 128, 128, 3, True), (17, 9, 128, 256, 3, False), (130, 70, 64, 64, 3, False),
    (19, 23, 256, 256, 1, False), (19, 23, 256, 65, 1, False)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_bf16x3(ops, H, W, cin, cout, k, pool, dtype):
    gen = torch.Generator().manual_seed(H * W + cin + cout)
    x = torch.relu(torch.randn(H, W, cin, generator=gen)) * 2.0
    w = torch.randn(cout, cin, k, k, generator=gen) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=gen) * 0.1
    pk = ops.PackedConv(w, b, "cuda", dtype)
    xs = _split(x, dtype).cuda()
    xq = _join(xs).cpu()                                    # what the kernel actually sees (16 mantissa bits)
    ref = F.conv2d(xq.permute(2, 0, 1)[None].double(), w.double(), b.double(), padding=k // 2)[0]     # [cout,H,W]
    scale = float(ref.abs().max())
    if cout % 64 == 0:
        y = ops.conv_bf16x3(xs, pk, relu=True, pool=pool, out="split")
        r = torch.relu(ref)
        if pool:
            r = F.max_pool2d(r[None], 2, 2)[0]
        assert y.shape == (2, r.shape[1], r.shape[2], cout)
        got = _join(y).cpu().double().permute(2, 0, 1)
        assert float((got - r).abs().max()) < 1e-4 * scale, float((got - r).abs().max()) / scale
    if not pool:
        y32 = ops.conv_bf16x3(xs, pk, relu=False, out="nhwc")
        assert y32.shape == (H, W, cout)
        err = float((y32.cpu().double().permute(2, 0, 1) - ref).abs().max())
        assert err < 1e-4 * scale, err / scale
        yp = ops.conv_bf16x3(xs, pk, relu=False, out="planar")
        assert torch.equal(yp.permute(1, 2, 0), y32)


def test_conv1a_split(ops):
    gen = torch.Generator().manual_seed(5)
    img = torch.rand(1, 1, 45, 67, generator=gen)
    w = torch.randn(64, 1, 3, 3, generator=gen) * 0.3
    b = torch.randn(64, generator=gen) * 0.1
    ref = torch.relu(F.conv2d(img.double(), w.double(), b.double(), padding=1))[0].permute(1, 2, 0)
    for dtype, tol in ((torch.bfloat16, 2e-5), (torch.float16, 2e-6)):
        y = ops.sp_conv1a_relu_split(img.cuda(), w.cuda(), b.cuda(), dtype)
        assert y.dtype == dtype
        assert float((_join(y).cpu().double() - ref).abs().max()) < tol * float(ref.abs().max())


@pytest.mark.parametrize("mode,tol", [("bf16x3", 2e-4), ("f16x3", 1e-4)])   # f16x3 is at the f32 summation-order noise of the cuDNN reference
def test_superpoint_backbone_bf16x3_vs_f32(mode, tol):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from icepy4d_b200.matching.superpoint import SuperPointB200
    from icepy4d_b200 import synthetic
    st = weights.make_superpoint_state(1)
    img, _ = synthetic.stereo_pair(240, 328, seed=3, shift=(4, 2), channels=1)
    x = torch.from_numpy(img.astype(np.float32) / 255.0)[None, None].cuda()
    ref = SuperPointB200(st, conv_precision="f32", keypoint_threshold=1e-4, max_keypoints=512)
    tcm = SuperPointB200(st, conv_precision=mode, keypoint_threshold=1e-4, max_keypoints=512)
    l0, d0 = ref.backbone(x)
    l1, d1 = tcm.backbone(x)
    assert l0.shape == l1.shape and d0.shape == d1.shape
    assert float((l0 - l1).abs().max()) < tol * float(l0.abs().max())
    assert float((d0 - d1).abs().max()) < tol * float(d0.abs().max())
    f0, f1 = ref.detect(x), tcm.detect(x)
    from icepy4d_b200.matching.superpoint import sync_counts
    sync_counts(f0, f1)
    k0 = {tuple(p) for p in f0.keypoints[:f0.n].cpu().numpy().tolist()}
    k1 = {tuple(p) for p in f1.keypoints[:f1.n].cpu().numpy().tolist()}
    assert len(k0 & k1) / len(k0 | k1) > 0.99


@pytest.mark.parametrize("H,W,pool", [(16, 16, True), (37, 53, True), (38, 54, False), (130, 199, True), (401, 333, True)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fused_conv1a_conv1b_is_bit_identical_to_the_two_kernels(ops, H, W, pool, dtype):
    """i4d_sp_conv1ab_tc computes relu(conv1a) inside conv1b's operand producer (the activation never goes to HBM).  Same f32
    FMA order, same (hi, lo) split, same shared-memory layout as the TMA box load with its zero fill: the output planes must be
    BIT-identical to i4d_sp_conv1a_relu followed by i4d_conv_bf16x3_tc (superpoint.py:154-156), at image borders included."""
    sd = weights.make_superpoint_state(1)
    gen = torch.Generator().manual_seed(H * 7 + W)
    img = torch.rand(1, 1, H, W, generator=gen).cuda()
    w1a, b1a = sd["conv1a.weight"].float().cuda().contiguous(), sd["conv1a.bias"].float().cuda().contiguous()
    pk = ops.PackedConv(sd["conv1b.weight"], sd["conv1b.bias"], "cuda", dtype)
    ref = ops.conv_bf16x3(ops.sp_conv1a_relu_split(img, w1a, b1a, dtype), pk, pool=pool)
    got = ops.sp_conv1ab_fused(img, w1a, b1a, pk, pool=pool)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
