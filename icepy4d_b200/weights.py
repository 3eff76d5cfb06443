"""Seeded *structured* random weights for SuperPoint / SuperGlue / LightGlue.

There are no pretrained checkpoints in this environment (`.MISSING_LARGE_BLOBS`,
LightGlue weights are URL downloads).  PyTorch's default init is degenerate for this path (flat score
maps, collapsed descriptors, zero matches), so benchmarks and parity tests use the recipe of SURVEY.md
Appendix B.  The state dicts produced here use the *reference's own parameter names and shapes*
(thirdparty/SuperGlue/models/superpoint.py:125-140, superglue.py:219-243, LightGlue/lightglue/lightglue.py:339-372)
so that a real `.pth` checkpoint can be dropped in through the same loaders.

All draws come from a CPU `torch.Generator` → bit-identical on every machine with the same torch build.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

StateDict = Dict[str, torch.Tensor]

_SP_CONVS = [  # name, cin, cout, k   (superpoint.py:125-140)
    ("conv1a", 1, 64, 3), ("conv1b", 64, 64, 3),
    ("conv2a", 64, 64, 3), ("conv2b", 64, 64, 3),
    ("conv3a", 64, 128, 3), ("conv3b", 128, 128, 3),
    ("conv4a", 128, 128, 3), ("conv4b", 128, 128, 3),
    ("convPa", 128, 256, 3), ("convPb", 256, 65, 1),
    ("convDa", 128, 256, 3), ("convDb", 256, 256, 1),
]


def make_superpoint_state(seed: int = 1) -> StateDict:
    """Appendix B, SuperPoint: N(0, 2/fan_in) filters made zero-mean, bias 0, convPb ×6."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for name, cin, cout, k in _SP_CONVS:
        fan_in = cin * k * k
        w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_in)
        w = w - w.mean(dim=(1, 2, 3), keepdim=True)
        if name == "convPb":
            w = w * 6.0
        sd[f"{name}.weight"] = w.contiguous()
        sd[f"{name}.bias"] = torch.zeros(cout)
    return sd


def _conv1d(g, cout, cin, scale=1.0):
    return (torch.randn(cout, cin, 1, generator=g) * math.sqrt(1.0 / cin) * scale).contiguous()


def _bn(sd, prefix, c):
    sd[f"{prefix}.weight"] = torch.ones(c)
    sd[f"{prefix}.bias"] = torch.zeros(c)
    sd[f"{prefix}.running_mean"] = torch.zeros(c)
    sd[f"{prefix}.running_var"] = torch.ones(c)
    sd[f"{prefix}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def make_superglue_state(seed: int = 2, n_layers: int = 18) -> StateDict:
    """Appendix B, SuperGlue (keys follow superglue.py: kenc.encoder.*, gnn.layers.*, final_proj, bin_score)."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    # keypoint encoder MLP([3, 32, 64, 128, 256, 256]) — Sequential idx: conv 0,3,6,9,12 ; bn 1,4,7,10
    chans = [3, 32, 64, 128, 256, 256]
    for i in range(1, len(chans)):
        idx = (i - 1) * 3
        last = i == len(chans) - 1
        sd[f"kenc.encoder.{idx}.weight"] = _conv1d(g, chans[i], chans[i - 1], 0.05 if last else 1.0)
        sd[f"kenc.encoder.{idx}.bias"] = torch.zeros(chans[i])
        if not last:
            _bn(sd, f"kenc.encoder.{idx + 1}", chans[i])
    for l in range(n_layers):
        p = f"gnn.layers.{l}"
        sd[f"{p}.attn.merge.weight"] = _conv1d(g, 256, 256)
        sd[f"{p}.attn.merge.bias"] = torch.zeros(256)
        for j in range(3):
            sd[f"{p}.attn.proj.{j}.weight"] = _conv1d(g, 256, 256)
            sd[f"{p}.attn.proj.{j}.bias"] = torch.zeros(256)
        sd[f"{p}.mlp.0.weight"] = _conv1d(g, 512, 512)
        sd[f"{p}.mlp.0.bias"] = torch.zeros(512)
        _bn(sd, f"{p}.mlp.1", 512)
        sd[f"{p}.mlp.3.weight"] = _conv1d(g, 256, 512, 0.02)
        sd[f"{p}.mlp.3.bias"] = torch.zeros(256)
    sd["final_proj.weight"] = (torch.eye(256) * math.sqrt(16.0 * 25.0)).unsqueeze(-1).contiguous()
    sd["final_proj.bias"] = torch.zeros(256)
    sd["bin_score"] = torch.tensor(1.0)
    return sd


def _linear(g, cout, cin, scale=1.0):
    return (torch.randn(cout, cin, generator=g) * math.sqrt(1.0 / cin) * scale).contiguous()


def make_lightglue_state(seed: int = 3, n_layers: int = 9, confident_layers=(), prune_variant: bool = False) -> StateDict:
    """Appendix B, LightGlue (keys follow lightglue.py: posenc.Wr, transformers.i.{self_attn,cross_attn}.*,
    log_assignment.i.*, token_confidence.i.token.0.*).  `confident_layers`: layer indices whose
    token-confidence bias is set to +3 so that early-stop / pruning paths are exercised.  `prune_variant`
    widens the matchability logits (std ~2, bias -2) so that point pruning really drops points."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    sd["posenc.Wr.weight"] = _linear(g, 32, 2)
    for i in range(n_layers):
        s = f"transformers.{i}.self_attn"
        sd[f"{s}.Wqkv.weight"] = _linear(g, 768, 256)
        sd[f"{s}.Wqkv.bias"] = torch.zeros(768)
        sd[f"{s}.out_proj.weight"] = _linear(g, 256, 256)
        sd[f"{s}.out_proj.bias"] = torch.zeros(256)
        c = f"transformers.{i}.cross_attn"
        for nm in ("to_qk", "to_v", "to_out"):
            sd[f"{c}.{nm}.weight"] = _linear(g, 256, 256)
            sd[f"{c}.{nm}.bias"] = torch.zeros(256)
        for blk in (s, c):
            sd[f"{blk}.ffn.0.weight"] = _linear(g, 512, 512)
            sd[f"{blk}.ffn.0.bias"] = torch.zeros(512)
            sd[f"{blk}.ffn.1.weight"] = torch.ones(512)
            sd[f"{blk}.ffn.1.bias"] = torch.zeros(512)
            sd[f"{blk}.ffn.3.weight"] = _linear(g, 256, 512, 0.02)
            sd[f"{blk}.ffn.3.bias"] = torch.zeros(256)
    for i in range(n_layers):
        a = f"log_assignment.{i}"
        sd[f"{a}.matchability.weight"] = _linear(g, 1, 256, 32.0 if prune_variant else 0.1)
        sd[f"{a}.matchability.bias"] = torch.full((1,), -2.0 if prune_variant else 4.0)
        sd[f"{a}.final_proj.weight"] = (torch.eye(256) * math.sqrt(16.0 * 25.0)).contiguous()
        sd[f"{a}.final_proj.bias"] = torch.zeros(256)
    for i in range(n_layers - 1):
        t = f"token_confidence.{i}.token.0"
        sd[f"{t}.weight"] = _linear(g, 1, 256)
        sd[f"{t}.bias"] = torch.full((1,), 3.0 if i in confident_layers else 0.0)
    return sd
