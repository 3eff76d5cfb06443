"""Tensor-core (tcgen05 / TMEM / TMA) path: bf16-operand GEMM + flash attention wrappers and the SuperGlue / LightGlue
layer schedules built on them.  f32 master copy of the residual stream, bf16 activations between kernels, f32
accumulation everywhere (TMEM)."""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import ops

BF16 = torch.bfloat16


_ATTN_WS = {}


def _st():
    return N.current_stream()


def gemm_tc(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            out32: Optional[torch.Tensor] = None, out16: Optional[torch.Tensor] = None, alpha: float = 1.0, relu: bool = False):
    """out = alpha * A @ W^T + bias (+relu) (+residual f32); A [M,K], W [N,K] bf16 (column slices allowed)."""
    assert A.dtype == BF16 and W.dtype == BF16 and A.is_cuda and W.is_cuda and A.stride(1) == 1 and W.stride(1) == 1
    M, K = A.shape
    Nn = W.shape[0]
    assert W.shape[1] == K
    assert out32 is not None or out16 is not None
    ld = lambda t, n: t.stride(0) if t.shape[0] > 1 else max(n, t.stride(0))
    N.call("i4d_gemm_bf16_tc", A, ld(A, K), W, ld(W, K), bias, residual, 0 if residual is None else ld(residual, Nn),
           out32, 0 if out32 is None else ld(out32, Nn), out16, 0 if out16 is None else ld(out16, Nn), M, Nn, K, float(alpha),
           int(relu), _st())
    return out32 if out32 is not None else out16


def gemm_tc_rotary(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, cs: torch.Tensor, rot_cols: int, out16: torch.Tensor):
    """out16 = rotary(A @ W^T + bias) on the first rot_cols output columns (q | k of LightGlue's fused QKV projection, heads of
    64 columns), plain projection on the rest; cs [M,64] f32 from lg_posenc.  One kernel instead of GEMM -> f32 -> rotary pass."""
    assert A.dtype == BF16 and W.dtype == BF16 and out16.dtype == BF16 and A.stride(1) == 1 and W.stride(1) == 1
    assert cs.dtype == torch.float32 and cs.is_contiguous() and cs.shape == (A.shape[0], 64)
    M, K = A.shape
    N.call("i4d_gemm_bf16_tc_rotary", A, A.stride(0), W, W.stride(0), bias, cs, int(rot_cols), out16, out16.stride(0), M,
           W.shape[0], K, _st())
    return out16


def attention_tc(X: torch.Tensor, problems: Sequence[Tuple[int, int, int, int]], out: torch.Tensor, q_col: int, k_col: int,
                 v_col: int, heads: int = 4, scale: float = 0.125, key_counts: Optional[torch.Tensor] = None):
    """X [rows, ld] bf16 holding Q/K/V column blocks; problems = [(q_row0, nq, k_row0, nk), ...]; out [rows, >=64*heads] bf16.
    key_counts (optional, device int32 [len(problems)]): how many of each problem's nk keys are real — the rest is masked on the
    device, so a launch (or a captured graph) with bucketed sizes serves every keypoint count of the bucket."""
    assert key_counts is None or (key_counts.dtype == torch.int32 and key_counts.is_cuda and key_counts.numel() >= len(problems))
    assert X.dtype == BF16 and out.dtype == BF16 and X.is_contiguous() and out.stride(1) == 1
    pr = np.ascontiguousarray(np.asarray(problems, dtype=np.int32).reshape(-1))
    key = (X.device.index, torch.cuda.current_stream().cuda_stream)      # one split-merge scratch per device and stream
    ws = _ATTN_WS.get(key)
    if ws is None:
        ws = _ATTN_WS[key] = torch.empty(N.lib().i4d_attention_workspace_bytes(), device=X.device, dtype=torch.uint8)
    N.call("i4d_attention_bf16_tc", X, X.shape[0], X.shape[1], int(q_col), int(k_col), int(v_col), int(heads), pr, len(problems),
           key_counts, float(scale), out, out.stride(0), ws, ws.numel(), _st())
    return out


def to_bf16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    assert x.dtype == torch.float32 and out.dtype == BF16 and x.stride(1) == 1 and out.stride(1) == 1
    N.call("i4d_f32_to_bf16", x, x.stride(0), out, out.stride(0), x.shape[0], x.shape[1], _st())
    return out


def split3_bf16(x: torch.Tensor, out: torch.Tensor, weight: bool = False) -> torch.Tensor:
    """x [n,c] f32 -> out [n,3c] bf16 = [hi | hi | lo] (activations) or [hi | lo | hi] (weights): operands of the three-product
    split-precision GEMM (one gemm_tc call of depth 3c)."""
    assert x.dtype == torch.float32 and out.dtype == BF16 and x.stride(1) == 1 and out.stride(1) == 1
    assert out.shape == (x.shape[0], 3 * x.shape[1])
    N.call("i4d_f32_split3_bf16", x, x.stride(0), out, out.stride(0), x.shape[0], x.shape[1], int(weight), _st())
    return out


def rotary_cast_bf16(qkv32: torch.Tensor, cs: torch.Tensor, out16: torch.Tensor) -> torch.Tensor:
    """qkv32 [n,768] f32, cs [n,64] -> out16 [n,768] bf16 with the rotary embedding applied to q and k (one pass)."""
    assert qkv32.dtype == torch.float32 and out16.dtype == BF16 and cs.dtype == torch.float32 and cs.is_contiguous()
    assert qkv32.shape[1] == 768 and out16.shape == qkv32.shape and cs.shape == (qkv32.shape[0], 64)
    N.call("i4d_lg_rotary_cast_bf16", qkv32, qkv32.stride(0), cs, qkv32.shape[0], out16, out16.stride(0), _st())
    return out16


def layernorm_gelu_bf16(x32: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out16: torch.Tensor, eps: float = 1e-5):
    """x32 [n,512] f32 -> GELU(LayerNorm(x32)) as bf16 (one pass)."""
    assert x32.dtype == torch.float32 and out16.dtype == BF16 and x32.shape == out16.shape and x32.shape[1] == 512
    N.call("i4d_layernorm_gelu_bf16", x32, x32.stride(0), gamma, beta, x32.shape[0], 512, float(eps), out16, out16.stride(0), _st())
    return out16


N._PER_CALL.update({"i4d_f32_split3_bf16": 1, "i4d_gemm_bf16_tc": 1, "i4d_gemm_bf16_tc_rotary": 1, "i4d_attention_bf16_tc": 1, "i4d_f32_to_bf16": 1, "i4d_lg_rotary_cast_bf16": 1,
                    "i4d_layernorm_gelu_bf16": 1})



class _CountsRing:
    """Real keypoint counts of a bucketed graph replay, host -> device without a stream synchronisation: a ring of pinned int32[4]
    staging buffers (a blocking copy from pageable memory would make the host wait for everything queued on the stream).  A slot is
    reused 32 replays later; the host can never be that far ahead of the device (every tile pair reads its keypoint counts back)."""

    def __init__(self, slots: int = 32):
        self.bufs = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(slots)]
        self.i = 0

    def push(self, dst: torch.Tensor, a: int, b: int):
        h = self.bufs[self.i]
        self.i = (self.i + 1) % len(self.bufs)
        h[0], h[1], h[2], h[3] = a, b, b, a
        dst.copy_(h, non_blocking=True)


class SuperGlueTensorCore:
    """GNN (18 layers) + final projection + score matrix of SuperGlue on the tensor-core path (superglue.py:131-149,276-280)."""

    def __init__(self, w, device):
        self.dev = torch.device(device)
        h = lambda t: t.to(BF16).contiguous()
        # `merge` (superglue.py:107,115) is linear and feeds only the first MLP layer: fold it into that layer's weights,
        #   W1 [x | merge(a)] + b1 = [W1x | W1m Wm] [x | a] + (b1 + W1m bm),   so the attention output goes straight into the
        # [x | message] buffer and one GEMM per layer disappears.
        self.layers = []
        for L in w.layers:
            w1x, w1m = L["w1"][:, :256], L["w1"][:, 256:]
            w1f = torch.cat([w1x, w1m @ L["wm"]], 1)
            b1f = L["b1"] + w1m @ L["bm"]
            self.layers.append({"wqkv": h(L["wqkv"]), "bqkv": L["bqkv"], "w1": h(w1f), "b1": b1f.contiguous(),
                                "w2": h(L["w2"]), "b2": L["b2"]})
        self.wf, self.bf = h(w.wf), w.bf
        # keypoint encoder (superglue.py:51-61,67-78): the two narrow layers (3 -> 32 -> 64) stay on the f32 SIMT GEMM, the three
        # wide ones (64 -> 128 -> 256 -> 256, 97 % of its flops) run on the tensor cores as three-product split-precision GEMMs
        # (weights pre-split [hi | lo | hi], activations split [hi | hi | lo] by one small kernel per layer): ~2^-16 relative error,
        # far inside the bf16 rounding the GNN applies to the encoder's output anyway
        self.kenc_f32 = list(w.kenc[:2])
        self.kenc_tc = []
        for wk, bk in w.kenc[2:]:
            w3 = torch.empty((wk.shape[0], 3 * wk.shape[1]), device=self.dev, dtype=BF16)
            split3_bf16(wk.contiguous(), w3, weight=True)
            self.kenc_tc.append((w3, bk))
        self.split_kenc = os.environ.get("I4D_SG_KENC_F32", "0") != "1"       # cross-check switch: all five layers on the f32 GEMM
        self._buf = {}
        self._graphs = {}
        self._static_buf = None
        self._counts_ring = _CountsRing()

    def encode(self, kpts, scores, desc, height, width) -> torch.Tensor:
        """desc + kenc([normalised kpts, score]) -> [n,256] f32 (superglue.py:67-78, 250-254)"""
        x = ops.sg_kenc_input(kpts, scores, float(width), float(height))
        for wk, bk in self.kenc_f32:
            x = ops.gemm_f32(x, wk, bk, relu=True)
        n = x.shape[0]
        for i, (w3, bk) in enumerate(self.kenc_tc):
            last = i == len(self.kenc_tc) - 1
            a3 = split3_bf16(x, torch.empty((n, 3 * x.shape[1]), device=x.device, dtype=BF16))
            x = gemm_tc(a3, w3, bk, residual=desc if last else None, out32=torch.empty((n, w3.shape[0]), device=x.device), relu=not last)
        return x

    def _buffers(self, nt):
        if self._buf.get("nt") != nt:
            self._buf = self._alloc(nt)
        return self._buf

    def _alloc(self, nt):
        d = self.dev
        return {"nt": nt, "x32": torch.empty((nt, 256), device=d), "xm": torch.empty((nt, 512), device=d, dtype=BF16),
                "qkv": torch.empty((nt, 768), device=d, dtype=BF16), "att": torch.empty((nt, 256), device=d, dtype=BF16),
                "hid": torch.empty((nt, 512), device=d, dtype=BF16), "md": torch.empty((nt, 256), device=d, dtype=BF16)}

    def _schedule(self, b, n0: int, n1: int, scores: torch.Tensor, collect=None, counts=None) -> torch.Tensor:
        """The 18-layer GNN + final projection + score GEMM on buffers `b`; b["x32"] already holds the encoded descriptors of
        image 0 in rows [0, n0) and of image 1 in rows [n0, n0 + n1).  counts (device int32 [4] = real n0, n1, n1, n0) masks the
        padding keys when n0 / n1 are bucket sizes."""
        nt = n0 + n1
        x32, xm, qkv, hid, md = b["x32"][:nt], b["xm"][:nt], b["qkv"][:nt], b["hid"][:nt], b["md"][:nt]
        to_bf16(x32, xm[:, :256])
        self_p = [(0, n0, 0, n0), (n0, n1, n0, n1)]
        cross_p = [(0, n0, n0, n1), (n0, n1, 0, n0)]
        kc_self, kc_cross = (None, None) if counts is None else (counts[:2], counts[2:])
        for l, L in enumerate(self.layers):
            gemm_tc(xm[:, :256], L["wqkv"], L["bqkv"], out16=qkv)
            cross = l % 2 == 1
            attention_tc(qkv, cross_p if cross else self_p, xm[:, 256:], 0, 256, 512, key_counts=kc_cross if cross else kc_self)
            gemm_tc(xm, L["w1"], L["b1"], out16=hid, relu=True)
            gemm_tc(hid, L["w2"], L["b2"], residual=x32, out32=x32, out16=xm[:, :256])
            if collect is not None:
                collect.append((x32[:n0].clone(), x32[n0:].clone()))
        gemm_tc(xm[:, :256], self.wf, self.bf, out16=md)
        gemm_tc(md[:n0], md[n0:], out32=scores, alpha=1.0 / 16.0)
        return scores

    def gnn_and_scores(self, d0: torch.Tensor, d1: torch.Tensor, collect=None) -> torch.Tensor:
        n0, n1 = d0.shape[0], d1.shape[0]
        nt = n0 + n1
        if collect is None and self.use_graphs and min(n0, n1) >= self.GRAPH_MIN:
            out = self._graphed(d0, d1)
            if out is not None:
                return out
        b = self._buffers(nt)
        b["x32"][:n0] = d0
        b["x32"][n0:] = d1
        scores = ops.padded_scores(n0, n1, self.dev)
        return self._schedule(b, n0, n1, scores, collect)

    # -- CUDA-graph replay of the schedule (75 launches per tile pair).  Graphs are keyed on SHAPE BUCKETS: both keypoint counts
    #    are rounded up to a multiple of GRAPH_BUCKET, the graph is captured at the bucket size on one set of static buffers
    #    (shared by all graphs, sized for the largest bucket seen), and the real counts live in a small device tensor the
    #    attention kernel reads (padding keys are masked; padding query rows cost a little arithmetic and are never read).  So
    #    tiles with uneven keypoint counts stay on the graph path.  The returned score matrix is a [n0, n1] view of the static
    #    buffer (row pitch = the bucket's n1): it is consumed (by the assignment kernels, same stream) before the next replay
    #    overwrites it.  The first call of a bucket runs eagerly (it also warms every kernel up); any capture failure disables
    #    graphs for good.
    use_graphs = os.environ.get("I4D_NO_GRAPHS", "0") != "1"
    MAX_GRAPHS = 16
    GRAPH_MIN = 1024
    GRAPH_BUCKET = 256

    def _static(self, nt: int, n_scores: int):
        st = self._static_buf
        if st is None or st["nt"] < nt or st["scores"].numel() < n_scores:
            torch.cuda.current_stream().synchronize()
            self._graphs.clear()                                              # graphs hold pointers into the old buffers
            nt = max(nt, 0 if st is None else st["nt"])
            n_scores = max(n_scores, 0 if st is None else st["scores"].numel())
            st = self._static_buf = self._alloc(nt)
            st["scores"] = torch.empty(n_scores, device=self.dev, dtype=torch.float32)
            st["counts"] = torch.zeros(4, device=self.dev, dtype=torch.int32)
        return st

    def _graphed(self, d0: torch.Tensor, d1: torch.Tensor):
        n0, n1 = d0.shape[0], d1.shape[0]
        B = self.GRAPH_BUCKET
        p0, p1 = -(-n0 // B) * B, -(-n1 // B) * B
        key = (p0, p1, torch.cuda.current_stream().cuda_stream)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.MAX_GRAPHS:
                torch.cuda.current_stream().synchronize()                     # its replay may still be running
                self._graphs.pop(next(iter(self._graphs)))                    # oldest bucket out
            self._graphs[key] = "seen"
            return None
        try:
            st = self._static(p0 + p1, p0 * p1)
            if key not in self._graphs:                                       # the static buffers grew: every graph was dropped
                self._graphs[key] = ent = "seen"
            scores = st["scores"][: p0 * p1].view(p0, p1)
            if ent == "seen":
                launches0 = N.LAUNCHES
                st["x32"][: p0 + p1].zero_()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._schedule(st, p0, p1, scores, counts=st["counts"])
                ent = self._graphs[key] = {"g": g, "launches": N.LAUNCHES - launches0}
                N.LAUNCHES = launches0           # capture launched nothing
            x32 = st["x32"]
            x32[:n0] = d0
            x32[p0: p0 + n1] = d1
            if n0 < p0:
                x32[n0:p0].zero_()               # padding rows restart from zero every replay (the residual stream is in place)
            if n1 < p1:
                x32[p0 + n1: p0 + p1].zero_()
            self._counts_ring.push(st["counts"], n0, n1)
            ent["g"].replay()
            N.LAUNCHES += ent["launches"]
            return scores[:n0, :n1]
        except Exception as err:                 # noqa: BLE001 - eager execution is always available
            import logging
            logging.getLogger(__name__).warning(f"CUDA-graph capture of the SuperGlue schedule failed ({err}); running eagerly")
            type(self).use_graphs = False
            self._graphs.clear()
            return None


class LightGlueTensorCore:
    """Self / cross / FFN blocks and the similarity matrix of LightGlue on the tensor-core path (lightglue.py:133-216,268-284).

    Both images are STACKED in one [m + n, .] set of buffers, so every GEMM / element-wise pass / attention launch of a layer
    covers both images (the weights are shared).  Per block the activations make one trip through HBM between two tensor-core
    kernels: the f32 master copy of the residual stream X [m+n,256] is read and written only by the GEMM epilogues (which also
    emit its bf16 shadow, the A operand of the next GEMM), rotary + cast and LayerNorm + GELU + cast are single passes."""

    def __init__(self, w, device):
        self.dev = torch.device(device)
        h = lambda t: t.to(BF16).contiguous()
        self.h = {}
        for L in w.layers:
            for k in ("wqkv", "wo", "wqkv_x", "wo_x", "w1_s", "w2_s", "w1_x", "w2_x"):
                self.h[id(L[k])] = h(L[k])
        for A in w.assign:
            self.h[id(A["wf"])] = h(A["wf"])
        self._buf = {}
        self._graphs = {}
        self._static_buf = None
        self._counts_ring = _CountsRing()

    def _w(self, t):
        return self.h[id(t)]

    def buffers(self, nt: int):
        """Persistent activation buffers for nt = m + n stacked keypoints (re-allocated only when nt grows)."""
        b = self._buf
        if b.get("cap", 0) < nt:
            if self._graphs:
                torch.cuda.current_stream().synchronize()
            self._graphs.clear()                                              # graphs hold pointers into the old buffers
            d = self.dev
            b = self._buf = {"cap": nt, "x16": torch.empty((nt, 512), device=d, dtype=BF16),
                             "qkv32": torch.empty((nt, 768), device=d), "qkv16": torch.empty((nt, 768), device=d, dtype=BF16),
                             "att": torch.empty((nt, 256), device=d, dtype=BF16), "hid32": torch.empty((nt, 512), device=d),
                             "hid16": torch.empty((nt, 512), device=d, dtype=BF16), "md": torch.empty((nt, 256), device=d, dtype=BF16)}
        return b

    def refresh_shadow(self, X: torch.Tensor):
        """bf16 shadow of the f32 residual stream (first layer, and after a pruning step re-packed X)."""
        to_bf16(X, self.buffers(X.shape[0])["x16"][: X.shape[0], :256])

    def _ffn(self, X, nt, L, tag, b):
        x16, hid32, hid16 = b["x16"][:nt], b["hid32"][:nt], b["hid16"][:nt]
        gemm_tc(x16, self._w(L[f"w1_{tag}"]), L[f"b1_{tag}"], out32=hid32)
        layernorm_gelu_bf16(hid32, L[f"g_{tag}"], L[f"be_{tag}"], hid16)
        gemm_tc(hid16, self._w(L[f"w2_{tag}"]), L[f"b2_{tag}"], residual=X, out32=X, out16=x16[:, :256])

    def layer(self, X: torch.Tensor, cs: torch.Tensor, m: int, n: int, L, counts: Optional[torch.Tensor] = None):
        """One LightGlue layer (self block on both images, cross block) on the stacked residual stream X [m+n,256] f32 (updated
        in place), cs [m+n,64] rotary tables.  The bf16 shadow of X must be current (`refresh_shadow`).  counts (device int32 [4] =
        real m, n, n, m) masks the padding keys when m / n are bucket sizes (CUDA-graph replay)."""
        nt = m + n
        b = self.buffers(nt)
        x16, qkv32, qkv16, att = b["x16"][:nt], b["qkv32"][:nt], b["qkv16"][:nt], b["att"][:nt]
        # ---- self block (lightglue.py:133-163): Wqkv -> rotary(q, k) -> attention -> out_proj -> FFN([x | message]) ----
        if self.fuse_rotary:
            gemm_tc_rotary(x16[:, :256], self._w(L["wqkv"]), L["bqkv"], cs, 512, qkv16)      # rotary(q, k) in the GEMM epilogue
        else:
            gemm_tc(x16[:, :256], self._w(L["wqkv"]), L["bqkv"], out32=qkv32)
            rotary_cast_bf16(qkv32, cs, qkv16)
        attention_tc(qkv16, [(0, m, 0, m), (m, n, m, n)], att, 0, 256, 512, key_counts=None if counts is None else counts[:2])
        gemm_tc(att, self._w(L["wo"]), L["bo"], out16=x16[:, 256:])
        self._ffn(X, nt, L, "s", b)
        # ---- cross block (lightglue.py:166-216): shared to_qk (q = k) and to_v, both directions in one attention launch ----
        p = qkv16[:, :512]                                                   # [qk | v]
        gemm_tc(x16[:, :256], self._w(L["wqkv_x"]), L["bqkv_x"], out16=p)
        attention_tc(qkv16, [(0, m, m, n), (m, n, 0, m)], att, 0, 0, 256, key_counts=None if counts is None else counts[2:])
        gemm_tc(att, self._w(L["wo_x"]), L["bo_x"], out16=x16[:, 256:])
        self._ffn(X, nt, L, "x", b)

    def similarity(self, X: torch.Tensor, m: int, n: int, A, out: Optional[torch.Tensor] = None):
        """sim = final_proj(x0) final_proj(x1)^T (lightglue.py:268-284; the 256^-1/4 scaling is folded into the weights)."""
        nt = m + n
        b = self.buffers(nt)
        md = b["md"][:nt]
        gemm_tc(b["x16"][:nt, :256], self._w(A["wf"]), A["bf"], out16=md)
        sim = ops.padded_scores(m, n, X.device) if out is None else out
        gemm_tc(md[:m], md[m:], out32=sim)
        return sim

    # -- CUDA-graph replay of the STATIC schedule (no early stop, no pruning: 9 layers + similarity, ~170 launches per tile pair),
    #    keyed on 256-keypoint shape buckets exactly like SuperGlueTensorCore._graphed: the layers run at the bucket size, the
    #    attention kernel masks the padding keys from a device tensor with the real counts, the caller gets the [m, n] corner of the
    #    bucket's similarity matrix and the final residual stream (for the matchability heads).
    use_graphs = os.environ.get("I4D_NO_GRAPHS", "0") != "1"
    fuse_rotary = os.environ.get("I4D_LG_NO_ROTARY_FUSION", "0") != "1"
    MAX_GRAPHS = 16
    GRAPH_MIN = 1024
    GRAPH_BUCKET = 256

    def _static(self, nt: int, n_sim: int):
        st = self._static_buf
        if st is None or st["nt"] < nt or st["sim"].numel() < n_sim:
            torch.cuda.current_stream().synchronize()
            self._graphs.clear()
            nt = max(nt, 0 if st is None else st["nt"])
            n_sim = max(n_sim, 0 if st is None else st["sim"].numel())
            d = self.dev
            st = self._static_buf = {"nt": nt, "X": torch.zeros((nt, 256), device=d), "cs": torch.zeros((nt, 64), device=d),
                                     "sim": torch.empty(n_sim, device=d), "counts": torch.zeros(4, device=d, dtype=torch.int32)}
        return st

    def graphed(self, desc0: torch.Tensor, desc1: torch.Tensor, cs0: torch.Tensor, cs1: torch.Tensor, layers, A):
        """-> (X0 [m,256], X1 [n,256] final residual streams, sim [m,n]) from a graph replay, or None (first call of a bucket,
        graphs disabled): the caller then runs the eager schedule."""
        m, n = desc0.shape[0], desc1.shape[0]
        if not self.use_graphs or min(m, n) < self.GRAPH_MIN:
            return None
        B = self.GRAPH_BUCKET
        p0, p1 = -(-m // B) * B, -(-n // B) * B
        key = (p0, p1, len(layers), id(A), torch.cuda.current_stream().cuda_stream)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.MAX_GRAPHS:
                torch.cuda.current_stream().synchronize()                     # its replay may still be running
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = "seen"           # first call of a bucket runs eagerly (it also warms every kernel up)
            return None
        try:
            self.buffers(p0 + p1)                # may grow (and drop every graph) BEFORE anything is captured
            st = self._static(p0 + p1, p0 * p1)
            if key not in self._graphs:
                self._graphs[key] = ent = "seen"
            X, cs = st["X"][: p0 + p1], st["cs"][: p0 + p1]
            sim = st["sim"][: p0 * p1].view(p0, p1)
            if ent == "seen":
                launches0 = N.LAUNCHES
                X.zero_()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.refresh_shadow(X)
                    for L in layers:
                        self.layer(X, cs, p0, p1, L, counts=st["counts"])
                    self.similarity(X, p0, p1, A, out=sim)
                ent = self._graphs[key] = {"g": g, "launches": N.LAUNCHES - launches0}
                N.LAUNCHES = launches0           # capture launched nothing
            X[:m] = desc0
            X[p0: p0 + n] = desc1
            cs[:m] = cs0
            cs[p0: p0 + n] = cs1
            if m < p0:
                X[m:p0].zero_()                  # padding rows restart from zero (the residual stream is updated in place)
                cs[m:p0].zero_()
            if n < p1:
                X[p0 + n:].zero_()
                cs[p0 + n:].zero_()
            self._counts_ring.push(st["counts"], m, n)
            ent["g"].replay()
            N.LAUNCHES += ent["launches"]
            return X[:m], X[p0: p0 + n], sim[:m, :n]
        except Exception as err:                 # noqa: BLE001 - eager execution is always available
            import logging
            logging.getLogger(__name__).warning(f"CUDA-graph capture of the LightGlue schedule failed ({err}); running eagerly")
            type(self).use_graphs = False
            self._graphs.clear()
            return None
