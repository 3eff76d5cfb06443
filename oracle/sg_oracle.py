"""TEST INFRASTRUCTURE — CPU oracle (torch fp32) for SuperGlue.  Never imported by the product.

Token-major restatement ([N, 256] rows instead of the reference's [256, N] Conv1d layout) of
/root/reference/src/icepy4d/thirdparty/SuperGlue/models/superglue.py:
  normalize_keypoints   :64-71      keypoint_encoder :51-61,74-84 (BatchNorm1d in eval mode)
  attention / MHA       :87-116     (head = channel % 4, because of `.view(b, dim, heads, n)` at :112)
  propagation + GNN     :119-149    final projection + scores :276-280
  log_optimal_transport :152-186    mutual-NN + threshold :288-298
Pinned against the reference itself by tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import math

import torch

BN_EPS = 1e-5


def normalize_keypoints(kpts: torch.Tensor, height: int, width: int) -> torch.Tensor:
    size = torch.tensor([float(width), float(height)])
    return (kpts - size / 2) / (size.max() * 0.7)


def _lin(x, sd, name):
    return x @ sd[f"{name}.weight"][:, :, 0].t() + sd[f"{name}.bias"]


def _bn(x, sd, name):
    inv = torch.rsqrt(sd[f"{name}.running_var"] + BN_EPS)
    return (x - sd[f"{name}.running_mean"]) * inv * sd[f"{name}.weight"] + sd[f"{name}.bias"]


def keypoint_encoder(kn: torch.Tensor, scores: torch.Tensor, sd: dict) -> torch.Tensor:
    x = torch.cat([kn, scores[:, None]], 1)
    for i in range(4):
        x = torch.relu(_bn(_lin(x, sd, f"kenc.encoder.{3 * i}"), sd, f"kenc.encoder.{3 * i + 1}"))
    return _lin(x, sd, "kenc.encoder.12")


def mha(x: torch.Tensor, src: torch.Tensor, sd: dict, p: str) -> torch.Tensor:
    n, m = x.shape[0], src.shape[0]
    q = _lin(x, sd, f"{p}.proj.0").view(n, 64, 4)       # channel c = d*4 + h
    k = _lin(src, sd, f"{p}.proj.1").view(m, 64, 4)
    v = _lin(src, sd, f"{p}.proj.2").view(m, 64, 4)
    s = torch.einsum("ndh,mdh->hnm", q, k) / 8.0
    pr = torch.softmax(s, -1)
    o = torch.einsum("hnm,mdh->ndh", pr, v).reshape(n, 256)
    return _lin(o, sd, f"{p}.merge")


def propagation(x, src, sd, l: int):
    p = f"gnn.layers.{l}"
    msg = mha(x, src, sd, f"{p}.attn")
    y = torch.cat([x, msg], 1)
    y = torch.relu(_bn(_lin(y, sd, f"{p}.mlp.0"), sd, f"{p}.mlp.1"))
    return _lin(y, sd, f"{p}.mlp.3")


def gnn(d0, d1, sd, n_layers=18, collect=None):
    for l in range(n_layers):
        cross = l % 2 == 1
        s0, s1 = (d1, d0) if cross else (d0, d1)
        e0, e1 = propagation(d0, s0, sd, l), propagation(d1, s1, sd, l)
        d0, d1 = d0 + e0, d1 + e1
        if collect is not None:
            collect.append((d0.clone(), d1.clone()))
    return d0, d1


def log_optimal_transport(scores: torch.Tensor, alpha: torch.Tensor, iters: int) -> torch.Tensor:
    m, n = scores.shape
    Z = scores.new_empty(m + 1, n + 1)
    Z[:m, :n] = scores
    Z[:m, n] = alpha
    Z[m, :] = alpha
    norm = -math.log(m + n)
    log_mu = torch.full((m + 1,), norm)
    log_mu[m] = math.log(n) + norm
    log_nu = torch.full((n + 1,), norm)
    log_nu[n] = math.log(m) + norm
    u, v = torch.zeros(m + 1), torch.zeros(n + 1)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v[None, :], 1)
        v = log_nu - torch.logsumexp(Z + u[:, None], 0)
    return Z + u[:, None] + v[None, :] - norm


def mutual_nn(P: torch.Tensor, thr: float):
    """P = log-assignment [M+1, N+1].  Returns matches0, matches1 (int64, -1 = none), mscores0, mscores1."""
    core = P[:-1, :-1]
    max0, max1 = core.max(1), core.max(0)
    i0, i1 = max0.indices, max1.indices
    mut0 = torch.arange(len(i0)) == i1[i0]
    mut1 = torch.arange(len(i1)) == i0[i1]
    ms0 = torch.where(mut0, max0.values.exp(), torch.zeros(()))
    ms1 = torch.where(mut1, ms0[i1], torch.zeros(()))
    v0 = mut0 & (ms0 > thr)
    v1 = mut1 & v0[i1]
    return torch.where(v0, i0, -1), torch.where(v1, i1, -1), ms0, ms1


def superglue(kpts0, sc0, desc0, kpts1, sc1, desc1, shape0, shape1, sd,
              iters=100, thr=0.2, n_layers=18, collect=None):
    """kpts [N,2] (x,y) f32; sc [N]; desc [256,N] (reference layout); shape = (H, W) of the image tensor."""
    d0 = desc0.t() + keypoint_encoder(normalize_keypoints(kpts0, *shape0), sc0, sd)
    d1 = desc1.t() + keypoint_encoder(normalize_keypoints(kpts1, *shape1), sc1, sd)
    if collect is not None:
        collect.append((d0.clone(), d1.clone()))
    d0, d1 = gnn(d0, d1, sd, n_layers, collect)
    m0, m1 = _lin(d0, sd, "final_proj"), _lin(d1, sd, "final_proj")
    scores = (m0 @ m1.t()) / 16.0
    P = log_optimal_transport(scores, sd["bin_score"], iters)
    a, b, c, d = mutual_nn(P, thr)
    return {"matches0": a, "matches1": b, "matching_scores0": c, "matching_scores1": d,
            "scores": scores, "log_assignment": P}
