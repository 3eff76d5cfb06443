"""SuperGlue on the device (token-major [N,256] activations, hand-written sm_100a kernels through the C ABI).

Mirrors thirdparty/SuperGlue/models/superglue.py:193-305 of the reference and accepts a state_dict with its
parameter names.  Host-side weight preparation (done once):
  * BatchNorm1d (eval) folded into the preceding Conv1d(k=1)                         (superglue.py:51-61)
  * the reference's head layout `view(b, 64, 4, n)` (channel = d*4 + h, superglue.py:112) is folded into a
    row permutation of the q/k/v projections and a column permutation of `merge`, so that the attention
    kernels see each head as 64 contiguous columns
  * q/k/v projections concatenated into one [768,256] matrix -> one GEMM per layer for both images

precision = "f32": CUDA-core kernels with exact f32 semantics (parity gate).
precision = "bf16": tcgen05 tensor-core GEMMs/attention with bf16 operands and f32 accumulation (throughput).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import ops

BN_EPS = 1e-5


def _fold_bn(w, b, sd, bn):
    g, beta = sd[f"{bn}.weight"], sd[f"{bn}.bias"]
    mu, var = sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"]
    s = g / torch.sqrt(var + BN_EPS)
    return w * s[:, None], (b - mu) * s + beta


def _head_perm():
    # new channel h*64 + d  <-  reference channel d*4 + h
    idx = torch.arange(256)
    h, d = idx // 64, idx % 64
    return d * 4 + h


class SuperGlueWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], device, n_layers: int = 18):
        f = lambda t: t.detach().to(torch.float32).cpu()
        sd = {k: f(v) if v.is_floating_point() else v for k, v in sd.items()}
        dev = torch.device(device)
        up = lambda t: t.contiguous().to(dev)
        self.kenc = []
        for i in range(5):
            w, b = sd[f"kenc.encoder.{3 * i}.weight"][:, :, 0], sd[f"kenc.encoder.{3 * i}.bias"]
            if i < 4:
                w, b = _fold_bn(w, b, sd, f"kenc.encoder.{3 * i + 1}")
            self.kenc.append((up(w), up(b)))
        perm = _head_perm()
        self.layers = []
        for l in range(n_layers):
            p = f"gnn.layers.{l}"
            wq = [sd[f"{p}.attn.proj.{j}.weight"][:, :, 0][perm] for j in range(3)]
            bq = [sd[f"{p}.attn.proj.{j}.bias"][perm] for j in range(3)]
            wm, bm = sd[f"{p}.attn.merge.weight"][:, :, 0][:, perm], sd[f"{p}.attn.merge.bias"]
            w1, b1 = _fold_bn(sd[f"{p}.mlp.0.weight"][:, :, 0], sd[f"{p}.mlp.0.bias"], sd, f"{p}.mlp.1")
            w2, b2 = sd[f"{p}.mlp.3.weight"][:, :, 0], sd[f"{p}.mlp.3.bias"]
            self.layers.append({"wqkv": up(torch.cat(wq, 0)), "bqkv": up(torch.cat(bq, 0)), "wm": up(wm), "bm": up(bm),
                                "w1": up(w1), "b1": up(b1), "w2": up(w2), "b2": up(b2)})
        self.wf, self.bf = up(sd["final_proj.weight"][:, :, 0]), up(sd["final_proj.bias"])
        self.bin_score = float(sd["bin_score"])
        self.n_layers = n_layers


class SuperGlueB200:
    def __init__(self, state_dict, device="cuda", sinkhorn_iterations: int = 100, match_threshold: float = 0.2,
                 precision: str = "f32", n_layers: int = 18):
        if not torch.cuda.is_available():
            raise RuntimeError("icepy4d_b200 needs a CUDA device (there is no CPU fallback)")
        assert precision in ("f32", "bf16")
        self.device = torch.device(device)
        self.w = SuperGlueWeights(state_dict, self.device, n_layers)
        self.iters, self.thr, self.precision = int(sinkhorn_iterations), float(match_threshold), precision
        self._ws = None
        if precision == "bf16":
            from .. import ops_tc
            self._tc = ops_tc.SuperGlueTensorCore(self.w, self.device)

    def encode(self, kpts, scores, desc, height, width) -> torch.Tensor:
        """desc + kenc([normalised kpts, score]) -> [n,256]"""
        x = ops.sg_kenc_input(kpts, scores, float(width), float(height))
        for i, (w, b) in enumerate(self.w.kenc):
            last = i == len(self.w.kenc) - 1
            x = ops.gemm_f32(x, w, b, residual=desc if last else None, relu=not last)
        return x

    def gnn_f32(self, d0: torch.Tensor, d1: torch.Tensor, collect=None) -> Tuple[torch.Tensor, torch.Tensor]:
        n0, n1 = d0.shape[0], d1.shape[0]
        nt = n0 + n1
        dev = self.device
        xm = torch.empty((nt, 512), device=dev, dtype=torch.float32)   # [x | message]
        xm[:n0, :256] = d0
        xm[n0:, :256] = d1
        qkv = torch.empty((nt, 768), device=dev, dtype=torch.float32)
        att = torch.empty((nt, 256), device=dev, dtype=torch.float32)
        hid = torch.empty((nt, 512), device=dev, dtype=torch.float32)
        x = xm[:, :256]
        for l, L in enumerate(self.w.layers):
            ops.gemm_f32(x, L["wqkv"], L["bqkv"], out=qkv)
            q, k, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
            cross = l % 2 == 1
            s0, s1 = (slice(n0, nt), slice(0, n0)) if cross else (slice(0, n0), slice(n0, nt))
            ops.attention_f32(q[:n0], k[s0], v[s0], att[:n0])
            ops.attention_f32(q[n0:], k[s1], v[s1], att[n0:])
            ops.gemm_f32(att, L["wm"], L["bm"], out=xm[:, 256:])
            ops.gemm_f32(xm, L["w1"], L["b1"], out=hid, relu=True)
            ops.gemm_f32(hid, L["w2"], L["b2"], residual=x, out=x)
            if collect is not None:
                collect.append((x[:n0].clone(), x[n0:].clone()))
        return x[:n0], x[n0:]

    def scores_f32(self, d0, d1) -> torch.Tensor:
        m0 = ops.gemm_f32(d0, self.w.wf, self.w.bf)
        m1 = ops.gemm_f32(d1, self.w.wf, self.w.bf)
        return ops.gemm_f32(m0, m1, alpha=1.0 / 16.0, out=ops.padded_scores(m0.shape[0], m1.shape[0], m0.device))

    def match(self, kpts0, sc0, desc0, shape0, kpts1, sc1, desc1, shape1, collect=None):
        """kpts [n,2], sc [n], desc [n,256] (token-major) on the device; shape = (H, W) of the image tensor.
        Returns matches0 [n0] i32, matches1 [n1] i32, mscores0, mscores1 (device tensors)."""
        n0, n1 = kpts0.shape[0], kpts1.shape[0]
        dev = self.device
        if n0 == 0 or n1 == 0:  # superglue.py:255-262
            return (torch.full((n0,), -1, device=dev, dtype=torch.int32), torch.full((n1,), -1, device=dev, dtype=torch.int32),
                    torch.zeros(n0, device=dev), torch.zeros(n1, device=dev))
        enc = self._tc.encode if self.precision == "bf16" and self._tc.split_kenc else self.encode
        d0 = enc(kpts0, sc0, desc0, *shape0)
        d1 = enc(kpts1, sc1, desc1, *shape1)
        if collect is not None:
            collect.append((d0.clone(), d1.clone()))
        if self.precision == "f32":
            d0, d1 = self.gnn_f32(d0, d1, collect)
            scores = self.scores_f32(d0, d1)
        else:
            scores = self._tc.gnn_and_scores(d0, d1, collect)
        if collect is not None:
            collect.append(scores)
        if self._ws is None or self._ws.M < n0 or self._ws.N < n1:
            self._ws = ops.AssignWorkspace(n0, n1, dev)
        return ops.sg_assign(scores, self.w.bin_score, self.iters, self.thr, self._ws)
