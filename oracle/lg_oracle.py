"""TEST INFRASTRUCTURE — CPU oracle (torch fp32) for LightGlue v0.1_arxiv.  Never imported by the product.

Functional restatement of /root/reference/src/icepy4d/thirdparty/LightGlue/lightglue/lightglue.py with the
semantics the reference has ON CPU (the only device it can run on here):
  normalize_keypoints :23-35   posenc/rotary :49-74   SelfBlock :133-162 (q/k/v interleaved stride 3 per head)
  CrossBlock :165-216 (CPU branch: one sim, two softmaxes)   TokenConfidence :77-89
  MatchAssignment / sigmoid_log_double_softmax :253-287       filter_matches :290-306
  _forward :436-556 incl. early stop (:491-494,571-579) and point pruning (:495-510, CPU threshold -1 => always on)
Pinned against the reference itself by tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def normalize_keypoints(kpts, size):
    size = torch.as_tensor(size, dtype=kpts.dtype)
    return (kpts - size / 2) / (size.max() / 2)


def posenc(kn, sd):
    proj = kn @ sd["posenc.Wr.weight"].t()                      # [N,32]
    return torch.cos(proj).repeat_interleave(2, -1), torch.sin(proj).repeat_interleave(2, -1)  # [N,64] each


def rotary(t, cs):
    """t [H,N,64]; pairs (2i, 2i+1) rotate: (x1,x2) -> (x1 c - x2 s, x2 c + x1 s)."""
    c, s = cs
    t2 = t.unflatten(-1, (-1, 2))
    rot = torch.stack((-t2[..., 1], t2[..., 0]), -1).flatten(-2)
    return t * c + rot * s


def _lin(x, sd, name):
    return x @ sd[f"{name}.weight"].t() + sd[f"{name}.bias"]


def _ffn(x, msg, sd, p):
    y = _lin(torch.cat([x, msg], -1), sd, f"{p}.ffn.0")
    y = F.layer_norm(y, (512,), sd[f"{p}.ffn.1.weight"], sd[f"{p}.ffn.1.bias"])
    return x + _lin(F.gelu(y), sd, f"{p}.ffn.3")


def self_block(x, enc, sd, p):
    n = x.shape[0]
    qkv = _lin(x, sd, f"{p}.Wqkv").view(n, 4, 64, 3).permute(1, 0, 2, 3)   # [H,N,64,3]
    q, k, v = rotary(qkv[..., 0], enc), rotary(qkv[..., 1], enc), qkv[..., 2]
    a = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    ctx = (a @ v).permute(1, 0, 2).reshape(n, 256)
    return _ffn(x, _lin(ctx, sd, f"{p}.out_proj"), sd, p)


def cross_block(x0, x1, sd, p):
    def heads(t):
        return t.view(t.shape[0], 4, 64).permute(1, 0, 2)

    s = 64 ** -0.25
    qk0, qk1 = heads(_lin(x0, sd, f"{p}.to_qk")) * s, heads(_lin(x1, sd, f"{p}.to_qk")) * s
    v0, v1 = heads(_lin(x0, sd, f"{p}.to_v")), heads(_lin(x1, sd, f"{p}.to_v"))
    sim = qk0 @ qk1.transpose(-1, -2)
    m0 = torch.softmax(sim, -1) @ v1
    m1 = torch.softmax(sim.transpose(-1, -2), -1) @ v0
    m0 = _lin(m0.permute(1, 0, 2).reshape(-1, 256), sd, f"{p}.to_out")
    m1 = _lin(m1.permute(1, 0, 2).reshape(-1, 256), sd, f"{p}.to_out")
    return _ffn(x0, m0, sd, p), _ffn(x1, m1, sd, p)


def log_assignment(d0, d1, sd, i):
    p = f"log_assignment.{i}"
    md0, md1 = _lin(d0, sd, f"{p}.final_proj") / 4.0, _lin(d1, sd, f"{p}.final_proj") / 4.0
    sim = md0 @ md1.t()
    z0, z1 = _lin(d0, sd, f"{p}.matchability"), _lin(d1, sd, f"{p}.matchability")
    m, n = sim.shape
    out = sim.new_zeros(m + 1, n + 1)
    out[:m, :n] = F.log_softmax(sim, 1) + F.log_softmax(sim, 0) + F.logsigmoid(z0) + F.logsigmoid(z1).t()
    out[:m, n] = F.logsigmoid(-z0[:, 0])
    out[m, :n] = F.logsigmoid(-z1[:, 0])
    return out, sim


def conf_threshold(i, n_layers=9):
    return float(np.clip(0.8 + 0.1 * np.exp(-4.0 * i / n_layers), 0, 1))


def lightglue(kpts0, desc0, size0, kpts1, desc1, size1, sd, n_layers=9, depth_conf=0.95,
              width_conf=0.99, filter_thr=0.1, collect=None):
    """kpts [N,2]; desc [N,256]; size = (W, H).  CPU semantics of LightGlue._forward for batch 1."""
    from .sg_oracle import mutual_nn

    m, n = len(kpts0), len(kpts1)
    e0, e1 = posenc(normalize_keypoints(kpts0, size0), sd), posenc(normalize_keypoints(kpts1, size1), sd)
    d0, d1 = desc0.clone(), desc1.clone()
    ind0, ind1 = torch.arange(m), torch.arange(n)
    prune0, prune1 = torch.ones(m, dtype=torch.long), torch.ones(n, dtype=torch.long)
    do_prune = width_conf > 0
    i = 0
    for i in range(n_layers):
        d0 = self_block(d0, e0, sd, f"transformers.{i}.self_attn")
        d1 = self_block(d1, e1, sd, f"transformers.{i}.self_attn")
        d0, d1 = cross_block(d0, d1, sd, f"transformers.{i}.cross_attn")
        if collect is not None:
            collect.append((d0.clone(), d1.clone()))
        if i == n_layers - 1:
            continue
        t0 = t1 = None
        if depth_conf > 0:
            w, b = sd[f"token_confidence.{i}.token.0.weight"], sd[f"token_confidence.{i}.token.0.bias"]
            t0, t1 = torch.sigmoid(d0 @ w.t() + b)[:, 0], torch.sigmoid(d1 @ w.t() + b)[:, 0]
            th = torch.tensor(conf_threshold(i, n_layers), dtype=torch.float32)
            ratio = 1.0 - (torch.cat([t0, t1]) < th).float().sum() / (m + n)
            if ratio > depth_conf:
                break
        if do_prune:
            th = torch.tensor(conf_threshold(i, n_layers), dtype=torch.float32)
            for side in (0, 1):
                d, t = (d0, t0) if side == 0 else (d1, t1)
                w = sd[f"log_assignment.{i}.matchability.weight"]
                b = sd[f"log_assignment.{i}.matchability.bias"]
                keep = torch.sigmoid(d @ w.t() + b)[:, 0] > (1 - width_conf)
                if t is not None:
                    keep = keep | (t <= th)
                idx = torch.where(keep)[0]
                if side == 0:
                    ind0, d0, e0 = ind0[idx], d0[idx], (e0[0][idx], e0[1][idx])
                    prune0[ind0] += 1
                else:
                    ind1, d1, e1 = ind1[idx], d1[idx], (e1[0][idx], e1[1][idx])
                    prune1[ind1] += 1
    P, sim = log_assignment(d0, d1, sd, i)
    a, b, c, d = mutual_nn(P, filter_thr)
    valid = a > -1
    matches = torch.stack([ind0[torch.where(valid)[0]], ind1[a[valid]]], -1)
    mscores = c[valid]
    if do_prune:
        m0 = torch.full((m,), -1, dtype=torch.long)
        m1 = torch.full((n,), -1, dtype=torch.long)
        m0[ind0] = torch.where(a == -1, -1, ind1[a.clamp(min=0)])
        m1[ind1] = torch.where(b == -1, -1, ind0[b.clamp(min=0)])
        s0, s1 = torch.zeros(m), torch.zeros(n)
        s0[ind0], s1[ind1] = c, d
        a, b, c, d = m0, m1, s0, s1
    else:
        prune0 = torch.full((m,), n_layers, dtype=torch.float32)
        prune1 = torch.full((n,), n_layers, dtype=torch.float32)
    return {"matches0": a, "matches1": b, "matching_scores0": c, "matching_scores1": d, "stop": i + 1,
            "matches": matches, "scores": mscores, "prune0": prune0, "prune1": prune1,
            "log_assignment": P, "sim": sim}
