"""Thin torch-tensor wrappers over the C ABI (validation + output allocation; all compute is in the .so).

Every function takes CUDA tensors, launches on torch's current stream and returns CUDA tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N


def _chk(t: torch.Tensor, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _st():
    return N.current_stream()


# ---------------------------------------------------------------- SuperPoint post-processing
def sp_score_map(logits: torch.Tensor) -> torch.Tensor:
    """logits [65,h,w] -> scores [8h,8w]."""
    _chk(logits, name="logits")
    c, h, w = logits.shape
    assert c == 65
    out = torch.empty((8 * h, 8 * w), device=logits.device, dtype=torch.float32)
    N.call("i4d_sp_score_map", logits, h, w, out, _st())
    return out


def sp_conv1a_relu(image: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
    """image [1,1,H,W] f32 -> relu(conv1a) as a channels-last [1,64,H,W] tensor of `out_dtype`."""
    _chk(image, name="image")
    H, W = image.shape[-2:]
    out = torch.empty((1, H, W, 64), device=image.device, dtype=out_dtype)
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[out_dtype]
    N.call("i4d_sp_conv1a_relu", image, H, W, _chk(weight.reshape(64, 9), name="weight"), _chk(bias, name="bias"), out, code, _st())
    return out.permute(0, 3, 1, 2)          # logical NCHW view of channels-last memory (no copy)


def maxpool2x2_cl(x: torch.Tensor) -> torch.Tensor:
    """x: [1,C,H,W] channels-last (post-ReLU) -> [1,C,H//2,W//2] channels-last."""
    assert x.is_cuda and x.dim() == 4 and x.shape[0] == 1 and x.is_contiguous(memory_format=torch.channels_last)
    _, C, H, W = x.shape
    out = torch.empty((1, H // 2, W // 2, C), device=x.device, dtype=x.dtype)
    N.call("i4d_maxpool2x2_nhwc", x, H, W, C, out, x.element_size(), 1, _st())
    return out.permute(0, 3, 1, 2)


def sp_conv1a_relu_split(image: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """image [1,1,H,W] f32 -> relu(conv1a) as split 16-bit planes [2,H,W,64] (hi, lo), bfloat16 or float16."""
    _chk(image, name="image")
    H, W = image.shape[-2:]
    out = torch.empty((2, H, W, 64), device=image.device, dtype=dtype)
    N.call("i4d_sp_conv1a_relu", image, H, W, _chk(weight.reshape(64, 9), name="weight"), _chk(bias, name="bias"), out,
           3 if dtype == torch.bfloat16 else 4, _st())
    return out


class PackedConv:
    """Weights of one convolution in the layout i4d_conv_bf16x3_tc reads (see include/icepy4d_b200.h)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device, dtype=torch.bfloat16):
        cout, cin, k, k2 = weight.shape
        self.dtype = dtype
        assert k == k2 and k in (1, 3) and cin % 64 == 0
        self.cout, self.cin, self.ksize = cout, cin, k
        self.cout_pad = 64 if cout <= 64 else -(-cout // 128) * 128
        n_t = 64 if self.cout_pad == 64 else 128
        nt, kc, taps = self.cout_pad // n_t, cin // 64, k * k
        wp = torch.zeros((self.cout_pad, cin, k, k), dtype=torch.float32)
        wp[:cout] = weight.detach().float().cpu()
        hi = wp.to(dtype)
        lo = (wp - hi.float()).to(dtype)
        t = torch.stack([hi, lo]).reshape(2, nt, n_t, kc, 64, taps)          # [part, nt, r, kc, k, tap]
        self.w = t.permute(1, 5, 3, 0, 2, 4).contiguous().reshape(-1, 64).to(device)   # [nt, tap, kc, part, r, k]
        b = torch.zeros(self.cout_pad, dtype=torch.float32)
        b[:cout] = bias.detach().float().cpu()
        self.b = b.to(device)


def conv_bf16x3(x: torch.Tensor, pk: PackedConv, relu: bool = True, pool: bool = False, out: str = "split") -> torch.Tensor:
    """x: split bf16 planes [2,H,W,Cin].  out = "split" -> [2,Ho,Wo,cout] bf16 planes (2x2 max-pool fused when pool),
    "nhwc" -> f32 [H,W,cout], "planar" -> f32 [cout,H,W]."""
    _chk(x, pk.dtype, "x")
    assert x.dim() == 4 and x.shape[0] == 2 and x.shape[3] == pk.cin
    H, W = int(x.shape[1]), int(x.shape[2])
    y16 = y32 = None
    if out == "split":
        Ho, Wo = (H // 2, W // 2) if pool else (H, W)
        y16 = torch.empty((2, Ho, Wo, pk.cout_pad), device=x.device, dtype=pk.dtype)
    elif out == "nhwc":
        y32 = torch.empty((H, W, pk.cout), device=x.device, dtype=torch.float32)
    elif out == "planar":
        y32 = torch.empty((pk.cout, H, W), device=x.device, dtype=torch.float32)
    else:
        raise ValueError(out)
    N.call("i4d_conv_bf16x3_tc", x[0], x[1], H, W, pk.cin, pk.w, pk.b, pk.cout_pad, pk.cout, pk.ksize, int(relu), int(pool),
           y16[0] if y16 is not None else None, y16[1] if y16 is not None else None, y32, pk.cout, int(out == "planar"),
           1 if pk.dtype == torch.bfloat16 else 0, _st())
    return y16 if y16 is not None else y32


def sp_conv1ab_fused(image: torch.Tensor, w1a: torch.Tensor, b1a: torch.Tensor, pk: PackedConv, pool: bool = True) -> torch.Tensor:
    """image [1,1,H,W] f32 -> relu(conv1b(relu(conv1a(image)))) as split planes [2,Ho,Wo,64] (2x2 max-pool fused when pool):
    one kernel, conv1a's activation never touches HBM (i4d_sp_conv1ab_tc)."""
    _chk(image, name="image")
    assert pk.cin == 64 and pk.cout == 64 and pk.ksize == 3
    H, W = int(image.shape[-2]), int(image.shape[-1])
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    y = torch.empty((2, Ho, Wo, 64), device=image.device, dtype=pk.dtype)
    N.call("i4d_sp_conv1ab_tc", image, H, W, _chk(w1a.reshape(64, 9), name="w1a"), _chk(b1a, name="b1a"), pk.w, pk.b, int(pool),
           y[0], y[1], 1 if pk.dtype == torch.bfloat16 else 0, _st())
    return y


class KeypointWorkspace:
    """Reusable device buffers for candidate compaction + top-k of one score map size."""

    def __init__(self, H: int, W: int, device, cand_cap: Optional[int] = None):
        self.H, self.W = H, W
        cap = int(cand_cap or max(65536, (H * W) // 4))
        self.cand_cap = 1 << (cap - 1).bit_length()   # power of two: the large-result sort pads to one
        self.keys = torch.empty(self.cand_cap, device=device, dtype=torch.int64)
        self.spill = torch.empty(2 * self.cand_cap + 4096, device=device, dtype=torch.int64)
        self.count = torch.zeros(1, device=device, dtype=torch.int32)


def sp_keypoints(scores: torch.Tensor, nms_radius: int, thr: float, border: int, k: int,
                 ws: Optional[KeypointWorkspace] = None, want_nms: bool = False, out_cap: Optional[int] = None):
    """scores [H,W] -> (kpts [cap,2] f32 (x,y), kscores [cap], n_dev int32[1], nms map or None).
    Only the first n rows are valid; n stays on the device (no sync here)."""
    _chk(scores, name="scores")
    H, W = scores.shape
    ws = ws or KeypointWorkspace(H, W, scores.device)
    nms = torch.empty_like(scores) if want_nms else None
    N.call("i4d_sp_nms_candidates", scores, H, W, int(nms_radius), float(thr), int(border), ws.keys, ws.cand_cap,
           ws.count, nms, _st())
    cap = int(out_cap if out_cap is not None else (k if k >= 0 else ws.cand_cap))
    kpts = torch.empty((cap, 2), device=scores.device, dtype=torch.float32)
    ksc = torch.empty((cap,), device=scores.device, dtype=torch.float32)
    n_dev = torch.zeros(1, device=scores.device, dtype=torch.int32)
    N.call("i4d_sp_select_topk", ws.keys, ws.count, ws.cand_cap, int(k), W, kpts, ksc, cap, n_dev, ws.spill, _st())
    return kpts, ksc, n_dev, nms


def sp_sample_descriptors(desc_hwc: torch.Tensor, kpts: torch.Tensor, n_dev: Optional[torch.Tensor] = None,
                          n_max: Optional[int] = None) -> torch.Tensor:
    """desc_hwc [h,w,256], kpts [n,2] -> [n,256] unit-norm rows."""
    _chk(desc_hwc, name="desc_hwc")
    _chk(kpts, name="kpts")
    h, w, c = desc_hwc.shape
    assert c == 256
    n_max = int(kpts.shape[0] if n_max is None else n_max)
    out = torch.zeros((n_max, 256), device=kpts.device, dtype=torch.float32)
    N.call("i4d_sp_sample_descriptors", desc_hwc, h, w, kpts, n_dev, n_max, out, _st())
    return out


# ---------------------------------------------------------------- dense f32
def gemm_f32(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, alpha: float = 1.0, relu: bool = False) -> torch.Tensor:
    """out[M,N] = alpha * A[M,K] @ W[N,K]^T + bias (+relu) (+residual).  A/W/out/residual may be column slices
    of wider row-major buffers (stride(1) == 1)."""
    for t, nm in ((A, "A"), (W, "W")):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(1) != 1:
            raise ValueError(f"{nm} must be a CUDA f32 matrix with unit column stride")
    M, K = A.shape
    Nn, K2 = W.shape
    assert K == K2
    if out is None:
        out = torch.empty((M, Nn), device=A.device, dtype=torch.float32)
    assert out.shape == (M, Nn) and out.stride(1) == 1
    ldr = 0
    if residual is not None:
        assert residual.shape == (M, Nn) and residual.stride(1) == 1
        ldr = residual.stride(0)
    N.call("i4d_gemm_f32", A, A.stride(0) if M > 1 else max(K, A.stride(0)), W, W.stride(0) if Nn > 1 else max(K, W.stride(0)),
           bias, residual, ldr, out, out.stride(0) if M > 1 else max(Nn, out.stride(0)), M, Nn, K, float(alpha), int(relu), _st())
    return out


def attention_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, heads: int = 4,
                  scale: float = 0.125) -> torch.Tensor:
    """q [Nq, heads*64], k/v [Nk, heads*64] (may be column slices), out [Nq, heads*64]."""
    for t in (q, k, v, out):
        assert t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1
    N.call("i4d_attention_f32", q, q.stride(0), k, k.stride(0), v, v.stride(0), out, out.stride(0), q.shape[0], k.shape[0],
           heads, float(scale), _st())
    return out


def layernorm_gelu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: Optional[torch.Tensor] = None,
                   eps: float = 1e-5) -> torch.Tensor:
    assert x.is_cuda and x.stride(1) == 1
    out = x if out is None else out
    N.call("i4d_layernorm_gelu", x, x.stride(0), gamma, beta, out, out.stride(0), x.shape[0], x.shape[1], float(eps), _st())
    return out


def lg_posenc(kpts: torch.Tensor, width: float, height: float, Wr: torch.Tensor) -> torch.Tensor:
    _chk(kpts, name="kpts")
    cs = torch.empty((kpts.shape[0], 64), device=kpts.device, dtype=torch.float32)
    N.call("i4d_lg_posenc", kpts, kpts.shape[0], float(width), float(height), _chk(Wr, name="Wr"), cs, _st())
    return cs


def lg_rotary_(x: torch.Tensor, cs: torch.Tensor, heads: int = 4) -> torch.Tensor:
    assert x.is_cuda and x.stride(1) == 1
    N.call("i4d_lg_rotary", x, x.stride(0), x.shape[0], heads, cs, _st())
    return x


def sg_kenc_input(kpts: torch.Tensor, scores: torch.Tensor, width: float, height: float) -> torch.Tensor:
    out = torch.empty((kpts.shape[0], 3), device=kpts.device, dtype=torch.float32)
    N.call("i4d_sg_kenc_input", _chk(kpts), _chk(scores), kpts.shape[0], float(width), float(height), out, _st())
    return out


# ---------------------------------------------------------------- assignment
class AssignWorkspace:
    def __init__(self, M: int, N_: int, device):
        self.M, self.N = M, N_
        nbytes = N.lib().i4d_assignment_workspace_bytes(M, N_)
        self.buf = torch.empty(nbytes, device=device, dtype=torch.uint8)
        self.nbytes = nbytes


def _ws(ws, M, N_, device):
    if ws is None or ws.M < M or ws.N < N_ or ws.nbytes < N.lib().i4d_assignment_workspace_bytes(M, N_):
        ws = AssignWorkspace(M, N_, device)
    return ws


def row_lse(S: torch.Tensor, scale: float = 1.0, coloff: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(S)
    out = torch.empty(S.shape[0], device=S.device, dtype=torch.float32)
    N.call("i4d_row_lse", S, S.shape[0], S.shape[1], float(scale), coloff, out, _st())
    return out


def col_lse(S: torch.Tensor, scale: float = 1.0, rowoff: Optional[torch.Tensor] = None, ws=None) -> torch.Tensor:
    _chk(S)
    ws = _ws(ws, S.shape[0], S.shape[1], S.device)
    out = torch.empty(S.shape[1], device=S.device, dtype=torch.float32)
    N.call("i4d_col_lse", S, S.shape[0], S.shape[1], float(scale), rowoff, out, ws.buf, ws.nbytes, _st())
    return out


def set_sinkhorn_mode(mode: int) -> None:
    """0 = fused persistent Sinkhorn when the shape allows (default), 1 = two-pass row/column kernels."""
    N.call("i4d_set_sinkhorn_mode", int(mode))


def _chk_pitched(t: torch.Tensor, name="matrix") -> int:
    """2-D f32 CUDA matrix with unit column stride; returns its row pitch in floats (`padded_scores` makes pitched ones)."""
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        raise ValueError(f"{name} must be a 2-D float32 CUDA tensor with unit column stride")
    return t.stride(0)


def padded_scores(M: int, N_: int, device) -> torch.Tensor:
    """[M, N] f32 view of a buffer whose row pitch is N rounded up to 4 floats: every row starts 16-byte aligned, which is
    what the fused Sinkhorn kernel (bulk row copies) and the 128-bit loads of the row/column passes need for any N."""
    return torch.empty((M, (N_ + 3) & ~3), device=device, dtype=torch.float32)[:, :N_]


def sinkhorn(scores: torch.Tensor, bin_score: float, iters: int, ws=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> u [M+1], v [N+1].  For N % 4 != 0 a pitched `scores` has its pad columns overwritten (see i4d_sinkhorn)."""
    ld = _chk_pitched(scores, "scores")
    M, N_ = scores.shape
    ws = _ws(ws, M, N_, scores.device)
    u = torch.empty(M + 1, device=scores.device, dtype=torch.float32)
    v = torch.empty(N_ + 4, device=scores.device, dtype=torch.float32)
    N.call("i4d_sinkhorn", scores, M, N_, ld, float(bin_score), int(iters), u, v, ws.buf, ws.nbytes, _st())
    return u, v[: N_ + 1]


def sg_assign(scores: torch.Tensor, bin_score: float, iters: int, thr: float, ws=None):
    """-> matches0 [M] i32, matches1 [N] i32, mscores0 [M], mscores1 [N]"""
    ld = _chk_pitched(scores, "scores")
    M, N_ = scores.shape
    ws = _ws(ws, M, N_, scores.device)
    dev = scores.device
    m0 = torch.empty(M, device=dev, dtype=torch.int32)
    m1 = torch.empty(N_, device=dev, dtype=torch.int32)
    s0 = torch.empty(M, device=dev, dtype=torch.float32)
    s1 = torch.empty(N_, device=dev, dtype=torch.float32)
    u = torch.empty(M + 1, device=dev, dtype=torch.float32)
    v = torch.empty(N_ + 4, device=dev, dtype=torch.float32)
    N.call("i4d_sg_assign", scores, M, N_, ld, float(bin_score), int(iters), float(thr), m0, m1, s0, s1, u, v, ws.buf,
           ws.nbytes, _st())
    return m0, m1, s0, s1


def lg_assign(sim: torch.Tensor, z0: torch.Tensor, z1: torch.Tensor, thr: float, ws=None):
    ld = _chk_pitched(sim, "sim")
    M, N_ = sim.shape
    ws = _ws(ws, M, N_, sim.device)
    dev = sim.device
    m0 = torch.empty(M, device=dev, dtype=torch.int32)
    m1 = torch.empty(N_, device=dev, dtype=torch.int32)
    s0 = torch.empty(M, device=dev, dtype=torch.float32)
    s1 = torch.empty(N_, device=dev, dtype=torch.float32)
    N.call("i4d_lg_assign", sim, M, N_, ld, _chk(z0.reshape(-1)), _chk(z1.reshape(-1)), float(thr), m0, m1, s0, s1, ws.buf,
           ws.nbytes, _st())
    return m0, m1, s0, s1


# ---------------------------------------------------------------- geometry
def undistort_points(pts: torch.Tensor, K: np.ndarray, dist: np.ndarray) -> torch.Tensor:
    _chk(pts, name="pts")
    Kh = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    dh = np.ascontiguousarray(np.asarray(dist, dtype=np.float64).reshape(-1))
    if dh.size > 5 and np.any(dh[5:] != 0):
        raise ValueError("only the 5-parameter Brown model (k1,k2,p1,p2,k3) is supported")
    dh = np.ascontiguousarray(dh[:5])
    out = torch.empty_like(pts)
    N.call("i4d_undistort_points", pts, pts.shape[0], Kh, dh, int(dh.size), out, _st())
    return out


def interpolate_point_colors(points3d: torch.Tensor, image_u8: torch.Tensor, R, t, K, dist, convert_bgr2rgb: bool = True,
                             want_projections: bool = False):
    """points3d [n,3] f64 and image [H,W,C] u8 on the device -> colors [n,C] f64 (and projections [n,2] f32)."""
    assert points3d.is_cuda and points3d.dtype == torch.float64 and points3d.is_contiguous() and points3d.shape[1] == 3
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.is_contiguous() and image_u8.dim() == 3
    Rh = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(9))
    th = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(3))
    Kh = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    dh = np.ascontiguousarray(np.asarray(dist, dtype=np.float64).reshape(-1))
    if dh.size > 5 and np.any(dh[5:] != 0):
        raise ValueError("only the 5-parameter Brown model (k1,k2,p1,p2,k3) is supported")
    dh = np.ascontiguousarray(dh[:5])
    n = points3d.shape[0]
    H, W, C = image_u8.shape
    col = torch.empty((n, C), device=points3d.device, dtype=torch.float64)
    proj = torch.empty((n, 2), device=points3d.device, dtype=torch.float32) if want_projections else None
    N.call("i4d_interpolate_point_colors", points3d, n, Rh, th, Kh, dh, int(dh.size), image_u8, H, W, C, int(bool(convert_bgr2rgb)),
           col, proj, _st())
    return (col, proj) if want_projections else col


def triangulate_iterative_ls(u1: torch.Tensor, u2: torch.Tensor, P1: np.ndarray, P2: np.ndarray, tol: float = 3e-5):
    _chk(u1), _chk(u2)
    n = u1.shape[0]
    X = torch.empty((n, 3), device=u1.device, dtype=torch.float64)
    st = torch.empty((n,), device=u1.device, dtype=torch.int32)
    P1h = np.ascontiguousarray(np.asarray(P1, dtype=np.float64).reshape(12))
    P2h = np.ascontiguousarray(np.asarray(P2, dtype=np.float64).reshape(12))
    N.call("i4d_triangulate_iterative_ls", u1, u2, n, P1h, P2h, float(tol), X, st, _st())
    return X, st


def triangulate_dlt(x1: torch.Tensor, x2: torch.Tensor, P1: np.ndarray, P2: np.ndarray) -> torch.Tensor:
    _chk(x1), _chk(x2)
    n = x1.shape[0]
    X = torch.empty((n, 3), device=x1.device, dtype=torch.float64)
    P1h = np.ascontiguousarray(np.asarray(P1, dtype=np.float64).reshape(12))
    P2h = np.ascontiguousarray(np.asarray(P2, dtype=np.float64).reshape(12))
    N.call("i4d_triangulate_dlt", x1, x2, n, P1h, P2h, X, _st())
    return X


MAGSAC_SIGMA_MAX = 4.5 / 3.64     # cut-off k * sigma_max = 4.5 px: fitted to OpenCV 4.13's USAC_MAGSAC output (scripts/magsac_probe.py)


def fundamental_ransac(x0: torch.Tensor, x1: torch.Tensor, threshold: float = 0.5, confidence: float = 0.999,
                       max_iters: int = 100000, seed: int = 0, sigma_max: float = MAGSAC_SIGMA_MAX, polish_iters: int = 64,
                       polish_mode: int = 0, ws: Optional[torch.Tensor] = None):
    """x0, x1 [n,2] f32 -> (F [9] f64 device, mask [n] u8 device, n_inliers int32[1] device).
    polish_mode 0 = MAGSAC++ weights (cv2 USAC_MAGSAC), 1 = least squares on the inliers at `threshold` (LO-RANSAC).
    F is NaN and the mask all zeros when no valid model exists (degenerate input), like cv2's (None, zeros)."""
    _chk(x0, name="x0"), _chk(x1, name="x1")
    n = x0.shape[0]
    dev = x0.device
    nbytes = N.lib().i4d_fundamental_workspace_bytes()
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    F = torch.zeros(9, device=dev, dtype=torch.float64)
    mask = torch.zeros(n, device=dev, dtype=torch.uint8)
    n_inl = torch.zeros(1, device=dev, dtype=torch.int32)
    N.call("i4d_fundamental_ransac", x0, x1, n, float(threshold), float(confidence), int(max_iters), int(seed) & 0xFFFFFFFF,
           float(sigma_max), int(polish_iters), int(polish_mode), F, mask, n_inl, ws, ws.numel(), _st())
    return F, mask, n_inl


def essential_ransac(xn0: torch.Tensor, xn1: torch.Tensor, threshold_norm: float, confidence: float = 0.9999, max_iters: int = 2048,
                     seed: int = 0):
    """Five-point RANSAC on K-normalised coordinates [n,2] f32 (n >= 5) -> (E [9] f64, n_inliers int32[1]) on the device."""
    _chk(xn0, name="xn0"), _chk(xn1, name="xn1")
    n = xn0.shape[0]
    dev = xn0.device
    ws = torch.empty(N.lib().i4d_essential_workspace_bytes(), device=dev, dtype=torch.uint8)
    E = torch.empty(9, device=dev, dtype=torch.float64)
    n_inl = torch.zeros(1, device=dev, dtype=torch.int32)
    N.call("i4d_essential_ransac", xn0, xn1, n, float(threshold_norm), float(confidence), int(max_iters), int(seed) & 0xFFFFFFFF, E,
           n_inl, ws, ws.numel(), _st())
    return E, n_inl


def essential_pose(E: torch.Tensor, xn0: torch.Tensor, xn1: torch.Tensor, threshold_norm: float, distance_threshold: float = 1e9):
    """E [9] f64 device, xn0 / xn1 [n,2] f32 normalised -> (E_proj [9], R [9], t [3] f64, mask [n] u8, n_good int32[2]) on the device."""
    _chk(xn0, name="xn0"), _chk(xn1, name="xn1")
    assert E.is_cuda and E.dtype == torch.float64 and E.numel() == 9
    n = xn0.shape[0]
    dev = xn0.device
    nbytes = N.lib().i4d_pose_workspace_bytes(n)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    Eo = torch.empty(9, device=dev, dtype=torch.float64)
    R = torch.empty(9, device=dev, dtype=torch.float64)
    t = torch.empty(3, device=dev, dtype=torch.float64)
    mask = torch.empty(n, device=dev, dtype=torch.uint8)
    n_good = torch.zeros(2, device=dev, dtype=torch.int32)
    N.call("i4d_essential_pose", E.contiguous(), xn0, xn1, n, float(threshold_norm), float(distance_threshold), Eo, R, t, mask,
           n_good, ws, ws.numel(), _st())
    return Eo, R, t, mask, n_good


def helmert_moments(v0: torch.Tensor, v1: torch.Tensor) -> torch.Tensor:
    """v0, v1 [n,3] f64 on the device -> [17] f64 (means, centred cross-covariance, sums of squares) on the device."""
    assert v0.is_cuda and v1.is_cuda and v0.dtype == torch.float64 and v1.dtype == torch.float64
    assert v0.shape == v1.shape and v0.dim() == 2 and v0.shape[1] == 3 and v0.shape[0] >= 1
    out = torch.empty(17, device=v0.device, dtype=torch.float64)
    N.call("i4d_helmert_moments", v0.contiguous(), v1.contiguous(), v0.shape[0], out, _st())
    return out


def apply_transform(X: torch.Tensor, T: np.ndarray) -> torch.Tensor:
    """X [n,3] f64 on the device, T 4x4 (host) -> dehomogenise(T [x; 1]) [n,3] f64 on the device."""
    assert X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.shape[1] == 3
    Th = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(16))
    X = X.contiguous()
    out = torch.empty_like(X)
    N.call("i4d_apply_transform", X, X.shape[0], Th, out, _st())
    return out


def tile_to_gray_f32(image_u8: torch.Tensor, x0: int, y0: int, tw: int, th: int, mode: int) -> torch.Tensor:
    """image [H,W,C] or [H,W] u8 on the device -> [1,1,th,tw] f32 network input."""
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.is_contiguous()
    H, W = image_u8.shape[:2]
    C = 1 if image_u8.dim() == 2 else image_u8.shape[2]
    out = torch.empty((1, 1, th, tw), device=image_u8.device, dtype=torch.float32)
    N.call("i4d_tile_to_gray_f32", image_u8, H, W, C, int(x0), int(y0), int(tw), int(th), int(mode), out, _st())
    return out


def tile_pair_counts(kp0: torch.Tensor, kp1: torch.Tensor, valid: Optional[torch.Tensor], scale: float, lims0, lims1) -> torch.Tensor:
    """PRESELECTION decision on the device: counts [T0,T1] i32 of valid matches strictly inside both tile rectangles
    (lims = sequences of (xmin, ymin, xmax, ymax), keypoints multiplied by `scale` first)."""
    _chk(kp0, name="kp0"), _chk(kp1, name="kp1")
    dev = kp0.device
    l0 = torch.tensor(np.asarray(lims0, dtype=np.float32).reshape(-1, 4), device=dev)
    l1 = torch.tensor(np.asarray(lims1, dtype=np.float32).reshape(-1, 4), device=dev)
    v = None if valid is None else valid.to(torch.uint8).contiguous()
    counts = torch.empty((l0.shape[0], l1.shape[0]), device=dev, dtype=torch.int32)
    N.call("i4d_tile_pair_counts", kp0, kp1, v, kp0.shape[0], float(scale), l0, l0.shape[0], l1, l1.shape[0], counts, _st())
    return counts


def pyr_down(image_u8: torch.Tensor) -> torch.Tensor:
    """cv2.pyrDown on a device u8 image [H,W] or [H,W,3] (bit-exact)."""
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.is_contiguous()
    H, W = image_u8.shape[:2]
    C = 1 if image_u8.dim() == 2 else image_u8.shape[2]
    shape = ((H + 1) // 2, (W + 1) // 2) + (() if image_u8.dim() == 2 else (C,))
    out = torch.empty(shape, device=image_u8.device, dtype=torch.uint8)
    N.call("i4d_pyr_down_u8", image_u8, H, W, C, out, _st())
    return out


def pyr_up(image_u8: torch.Tensor) -> torch.Tensor:
    """cv2.pyrUp on a device u8 image [H,W] or [H,W,3] (bit-exact)."""
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.is_contiguous()
    H, W = image_u8.shape[:2]
    C = 1 if image_u8.dim() == 2 else image_u8.shape[2]
    shape = (2 * H, 2 * W) + (() if image_u8.dim() == 2 else (C,))
    out = torch.empty(shape, device=image_u8.device, dtype=torch.uint8)
    N.call("i4d_pyr_up_u8", image_u8, H, W, C, out, _st())
    return out
