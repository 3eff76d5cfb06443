"""LightGlue v0.1_arxiv on the device (token-major activations, hand-written sm_100a kernels through the C ABI).

Mirrors thirdparty/LightGlue/lightglue/lightglue.py:309-585 of the reference with the semantics it has on CPU
(the only device the oracle can run on): point pruning is evaluated at every layer (`pruning_keypoint_thresholds
['cpu'] = -1`, lightglue.py:326-331), attention in f32.  Accepts a state_dict with the reference's parameter names.

Host-side weight preparation (once):
  * Wqkv rows permuted from the reference's interleaved layout (row = h*192 + d*3 + {q,k,v}; lightglue.py:155-156)
    to [q | k | v] x [head][dim] so every head is 64 contiguous columns;
  * cross-attention to_qk / to_v concatenated into one [512,256] projection;
  * the d**-0.25 scaling of MatchAssignment.final_proj folded into its weight and bias (lightglue.py:275-277).

The adaptive depth/width logic needs data-dependent control flow (one host read per layer, exactly as the reference's
`break` / `torch.where`); pass depth_confidence=-1, width_confidence=-1 for the static, sync-free mode.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .. import ops


def _qkv_perm():
    r = torch.arange(768)
    t, h, d = r // 256, (r % 256) // 64, r % 64
    return h * 192 + d * 3 + t


class LightGlueWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], device, n_layers: int = 9):
        sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items() if v.is_floating_point()}
        dev = torch.device(device)
        up = lambda t: t.contiguous().to(dev)
        self.Wr = up(sd["posenc.Wr.weight"])
        perm = _qkv_perm()
        self.layers = []
        for i in range(n_layers):
            s, c = f"transformers.{i}.self_attn", f"transformers.{i}.cross_attn"
            L = {"wqkv": up(sd[f"{s}.Wqkv.weight"][perm]), "bqkv": up(sd[f"{s}.Wqkv.bias"][perm]),
                 "wo": up(sd[f"{s}.out_proj.weight"]), "bo": up(sd[f"{s}.out_proj.bias"]),
                 "wqkv_x": up(torch.cat([sd[f"{c}.to_qk.weight"], sd[f"{c}.to_v.weight"]], 0)),
                 "bqkv_x": up(torch.cat([sd[f"{c}.to_qk.bias"], sd[f"{c}.to_v.bias"]], 0)),
                 "wo_x": up(sd[f"{c}.to_out.weight"]), "bo_x": up(sd[f"{c}.to_out.bias"])}
            for tag, blk in (("s", s), ("x", c)):
                L[f"w1_{tag}"], L[f"b1_{tag}"] = up(sd[f"{blk}.ffn.0.weight"]), up(sd[f"{blk}.ffn.0.bias"])
                L[f"g_{tag}"], L[f"be_{tag}"] = up(sd[f"{blk}.ffn.1.weight"]), up(sd[f"{blk}.ffn.1.bias"])
                L[f"w2_{tag}"], L[f"b2_{tag}"] = up(sd[f"{blk}.ffn.3.weight"]), up(sd[f"{blk}.ffn.3.bias"])
            self.layers.append(L)
        self.assign = []
        for i in range(n_layers):
            a = f"log_assignment.{i}"
            self.assign.append({"wf": up(sd[f"{a}.final_proj.weight"] / 4.0), "bf": up(sd[f"{a}.final_proj.bias"] / 4.0),
                                "wm": up(sd[f"{a}.matchability.weight"]), "bm": up(sd[f"{a}.matchability.bias"])})
        self.token = [{"w": up(sd[f"token_confidence.{i}.token.0.weight"]), "b": up(sd[f"token_confidence.{i}.token.0.bias"])}
                      for i in range(n_layers - 1)]
        self.n_layers = n_layers
        # lightglue.py:558-561 — float32 buffer of thresholds
        self.conf_thr = torch.tensor([float(np.clip(0.8 + 0.1 * np.exp(-4.0 * i / n_layers), 0, 1)) for i in range(n_layers)],
                                     dtype=torch.float32, device=dev)


class LightGlueB200:
    def __init__(self, state_dict, device="cuda", precision: str = "f32", depth_confidence: float = 0.95,
                 width_confidence: float = 0.99, filter_threshold: float = 0.1, n_layers: int = 9):
        if not torch.cuda.is_available():
            raise RuntimeError("icepy4d_b200 needs a CUDA device (there is no CPU fallback)")
        assert precision in ("f32", "bf16")
        self.device = torch.device(device)
        self.w = LightGlueWeights(state_dict, self.device, n_layers)
        self.precision = precision
        self.depth_confidence, self.width_confidence = float(depth_confidence), float(width_confidence)
        self.filter_threshold = float(filter_threshold)
        self._ws = None
        self._tc = None
        if precision == "bf16":
            from .. import ops_tc
            self._tc = ops_tc.LightGlueTensorCore(self.w, self.device)

    # ---- building blocks (f32 path) ----
    def _ffn(self, xm: torch.Tensor, L, tag: str):
        """xm [n,512] = [x | message]; x <- x + W2 gelu(LN(W1 [x|m]))   (lightglue.py:144-149,162)"""
        x = xm[:, :256]
        h = ops.gemm_f32(xm, L[f"w1_{tag}"], L[f"b1_{tag}"])
        ops.layernorm_gelu(h, L[f"g_{tag}"], L[f"be_{tag}"])
        ops.gemm_f32(h, L[f"w2_{tag}"], L[f"b2_{tag}"], residual=x, out=x)

    def _self_block(self, xm: torch.Tensor, cs: torch.Tensor, L):
        x = xm[:, :256]
        qkv = ops.gemm_f32(x, L["wqkv"], L["bqkv"])
        ops.lg_rotary_(qkv[:, :256], cs)
        ops.lg_rotary_(qkv[:, 256:512], cs)
        att = torch.empty((x.shape[0], 256), device=x.device, dtype=torch.float32)
        ops.attention_f32(qkv[:, :256], qkv[:, 256:512], qkv[:, 512:], att)
        ops.gemm_f32(att, L["wo"], L["bo"], out=xm[:, 256:])
        self._ffn(xm, L, "s")

    def _cross_block(self, xm0: torch.Tensor, xm1: torch.Tensor, L):
        p0 = ops.gemm_f32(xm0[:, :256], L["wqkv_x"], L["bqkv_x"])     # [m,512] = [qk | v]
        p1 = ops.gemm_f32(xm1[:, :256], L["wqkv_x"], L["bqkv_x"])
        a0 = torch.empty((xm0.shape[0], 256), device=xm0.device, dtype=torch.float32)
        a1 = torch.empty((xm1.shape[0], 256), device=xm1.device, dtype=torch.float32)
        ops.attention_f32(p0[:, :256], p1[:, :256], p1[:, 256:], a0)   # both directions share qk0, qk1 (lightglue.py:201-208)
        ops.attention_f32(p1[:, :256], p0[:, :256], p0[:, 256:], a1)
        ops.gemm_f32(a0, L["wo_x"], L["bo_x"], out=xm0[:, 256:])
        ops.gemm_f32(a1, L["wo_x"], L["bo_x"], out=xm1[:, 256:])
        self._ffn(xm0, L, "x")
        self._ffn(xm1, L, "x")

    def _lin1(self, x: torch.Tensor, w, b) -> torch.Tensor:
        return ops.gemm_f32(x, w, b)[:, 0]

    # ---- forward (lightglue.py:436-556, batch of 1) ----
    def match(self, kpts0, desc0, size0, kpts1, desc1, size1, collect=None):
        """kpts [n,2], desc [n,256] on the device; size = (W, H) of the image.  Returns a dict of device tensors."""
        dev = self.device
        m, n = kpts0.shape[0], kpts1.shape[0]
        W = self.w
        nl = W.n_layers
        if m == 0 or n == 0:
            e = lambda k, dt: torch.full((k,), -1, device=dev, dtype=dt)
            return {"matches0": e(m, torch.int32), "matches1": e(n, torch.int32), "matching_scores0": torch.zeros(m, device=dev),
                    "matching_scores1": torch.zeros(n, device=dev), "stop": 0, "matches": torch.zeros((0, 2), dtype=torch.int64, device=dev),
                    "scores": torch.zeros(0, device=dev), "prune0": torch.ones(m, device=dev), "prune1": torch.ones(n, device=dev)}
        cs0 = ops.lg_posenc(kpts0.contiguous(), float(size0[0]), float(size0[1]), W.Wr)
        cs1 = ops.lg_posenc(kpts1.contiguous(), float(size1[0]), float(size1[1]), W.Wr)
        tc = self._tc
        do_stop, do_prune = self.depth_confidence > 0, self.width_confidence > 0
        sim = None
        if tc is not None and not do_stop and not do_prune and collect is None:
            # static schedule: all layers + the similarity matrix replayed from a CUDA graph (keyed on 256-keypoint buckets)
            res = tc.graphed(desc0, desc1, cs0, cs1, W.layers[:nl], W.assign[nl - 1])
            if res is not None:
                x0, x1, sim = res
        if sim is not None:
            pass
        elif tc is not None:
            # tensor-core path: both images stacked in one f32 residual stream X [m+n,256] (+ its bf16 shadow inside `tc`)
            X = torch.empty((m + n, 256), device=dev, dtype=torch.float32)
            X[:m] = desc0
            X[m:] = desc1
            cs = torch.cat([cs0, cs1])
            tc.refresh_shadow(X)
            mc, nc = m, n                                                        # current (un-pruned) keypoint counts
            x0, x1 = X[:mc], X[mc:]
        else:
            xm0 = torch.empty((m, 512), device=dev, dtype=torch.float32)
            xm1 = torch.empty((n, 512), device=dev, dtype=torch.float32)
            xm0[:, :256] = desc0
            xm1[:, :256] = desc1
            x0, x1 = xm0[:, :256], xm1[:, :256]
        ind0, ind1 = torch.arange(m, device=dev), torch.arange(n, device=dev)
        prune0 = torch.ones(m, dtype=torch.int64, device=dev)
        prune1 = torch.ones(n, dtype=torch.int64, device=dev)
        i = nl - 1 if sim is not None else 0
        for i in (range(nl) if sim is None else ()):
            L = W.layers[i]
            if tc is not None:
                tc.layer(X, cs, mc, nc, L)
            else:
                self._self_block(xm0, cs0, L)
                self._self_block(xm1, cs1, L)
                self._cross_block(xm0, xm1, L)
            if collect is not None:
                collect.append((x0.clone(), x1.clone()))
            if i == nl - 1:
                continue
            t0 = t1 = None
            if do_stop:   # lightglue.py:491-494,571-579
                tk = W.token[i]
                t0 = torch.sigmoid(self._lin1(x0, tk["w"], tk["b"]))
                t1 = torch.sigmoid(self._lin1(x1, tk["w"], tk["b"]))
                unconf = (torch.cat([t0, t1]) < W.conf_thr[i]).float().sum()
                if float(1.0 - unconf / (m + n)) > self.depth_confidence:      # host read, as in the reference
                    break
            if do_prune:  # lightglue.py:495-510 (CPU semantics: evaluated at every layer)
                A = W.assign[i]
                pruned = False
                for side in (0, 1):
                    x, t = (x0, t0) if side == 0 else (x1, t1)
                    keep = torch.sigmoid(self._lin1(x, A["wm"], A["bm"])) > (1 - self.width_confidence)
                    if t is not None:
                        keep = keep | (t <= W.conf_thr[i])
                    idx = torch.nonzero(keep).squeeze(1)                          # host read (data-dependent size)
                    if idx.numel() != x.shape[0]:
                        pruned = True
                        if side == 0:
                            ind0, cs0 = ind0[idx], cs0.index_select(0, idx)
                            if tc is not None:
                                x0 = x0.index_select(0, idx)
                            else:
                                xm0 = xm0.index_select(0, idx)
                                x0 = xm0[:, :256]
                        else:
                            ind1, cs1 = ind1[idx], cs1.index_select(0, idx)
                            if tc is not None:
                                x1 = x1.index_select(0, idx)
                            else:
                                xm1 = xm1.index_select(0, idx)
                                x1 = xm1[:, :256]
                    if side == 0:
                        prune0[ind0] += 1
                    else:
                        prune1[ind1] += 1
                if pruned and tc is not None:                                    # re-pack the stacked stream and its shadow
                    mc, nc = x0.shape[0], x1.shape[0]
                    X = torch.cat([x0, x1])
                    cs = torch.cat([cs0, cs1])
                    x0, x1 = X[:mc], X[mc:]
                    if mc + nc > 0:
                        tc.refresh_shadow(X)
                if x0.shape[0] == 0 or x1.shape[0] == 0:
                    break
        A = W.assign[i]
        if x0.shape[0] == 0 or x1.shape[0] == 0:
            a = torch.full((x0.shape[0],), -1, device=dev, dtype=torch.int32)
            b = torch.full((x1.shape[0],), -1, device=dev, dtype=torch.int32)
            c, d = torch.zeros(x0.shape[0], device=dev), torch.zeros(x1.shape[0], device=dev)
        else:
            if sim is not None:
                pass                                                             # from the graph replay
            elif tc is not None:
                sim = tc.similarity(X, x0.shape[0], x1.shape[0], A)
            else:
                md0, md1 = ops.gemm_f32(x0, A["wf"], A["bf"]), ops.gemm_f32(x1, A["wf"], A["bf"])
                sim = ops.gemm_f32(md0, md1, out=ops.padded_scores(md0.shape[0], md1.shape[0], md0.device))
            z0, z1 = self._lin1(x0, A["wm"], A["bm"]).contiguous(), self._lin1(x1, A["wm"], A["bm"]).contiguous()
            if collect is not None:
                collect.append(sim)
            if self._ws is None or self._ws.M < sim.shape[0] or self._ws.N < sim.shape[1]:
                self._ws = ops.AssignWorkspace(sim.shape[0], sim.shape[1], dev)
            a, b, c, d = ops.lg_assign(sim, z0, z1, self.filter_threshold, self._ws)
        valid = a > -1
        vi = torch.nonzero(valid).squeeze(1)
        matches = torch.stack([ind0[vi], ind1[a[vi].long()]], -1)
        mscores = c[vi]
        if do_prune:   # lightglue.py:528-539
            m0 = torch.full((m,), -1, device=dev, dtype=torch.int32)
            m1 = torch.full((n,), -1, device=dev, dtype=torch.int32)
            m0[ind0] = torch.where(a == -1, a, ind1[a.clamp(min=0).long()].to(torch.int32))
            m1[ind1] = torch.where(b == -1, b, ind0[b.clamp(min=0).long()].to(torch.int32))
            s0, s1 = torch.zeros(m, device=dev), torch.zeros(n, device=dev)
            s0[ind0], s1[ind1] = c, d
            a, b, c, d = m0, m1, s0, s1
        else:
            prune0 = torch.full((m,), float(nl), device=dev)
            prune1 = torch.full((n,), float(nl), device=dev)
        return {"matches0": a, "matches1": b, "matching_scores0": c, "matching_scores1": d, "stop": i + 1,
                "matches": matches, "scores": mscores, "prune0": prune0, "prune1": prune1}
