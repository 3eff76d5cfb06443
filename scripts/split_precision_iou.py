"""Experiment: what operand precision does the SuperPoint backbone need for >=99 % end-to-end match IoU?

Emulates split-operand tensor-core convolutions with exact f32 cuDNN convolutions on pre-rounded operands
(products of bf16 values are exact in f32, accumulation is f32 either way):
  bf16x3   x = x1 + x2 (bf16 each), w likewise; x1*w1 + x2*w1 + x1*w2      (relative error ~2^-16)
  bf16x6   three-way split, six products                                    (~2^-24)
  tf32     operands truncated to 10 mantissa bits                           (~2^-11, what cuDNN TF32 does)
  mixed    tf32 everywhere except the named layers, which run in f32
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from icepy4d_b200 import weights
from icepy4d_b200.matching import SuperGlueMatcher, LightGlueMatcher, GeometricVerification, Quality, TileSelection
from icepy4d_b200.matching import superpoint as sp_mod

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/matchers.npz"))
def pairs(a, b): return {(float(p[0]), float(p[1]), float(q[0]), float(q[1])) for p, q in zip(a, b)}
ref = pairs(g["sg_mkpts0"], g["sg_mkpts1"]); refl = pairs(g["lg_mkpts0"], g["lg_mkpts1"])


def split_bf16(x, n):
    parts = []
    r = x
    for _ in range(n):
        p = r.to(torch.bfloat16).float()
        parts.append(p)
        r = r - p
    return parts


def split_f16(x, n):
    parts = []
    r = x
    for _ in range(n):
        p = r.to(torch.float16).float()
        parts.append(p)
        r = r - p
    return parts


def tf32_round(x):
    i = x.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF        # round-to-nearest on the 13 dropped bits
    return i.view(torch.float32)


MODE = {"kind": "f32", "f32_layers": ()}


def emu_conv(self, x, name, pad, relu=True):
    w, b = self.w[name]
    x = x.float(); w = w.float(); b = b.float()
    kind = MODE["kind"]
    if name in MODE["f32_layers"]:
        kind = "f32"
    if kind == "f32":
        y = F.conv2d(x, w, b, padding=pad)
    elif kind == "tf32":
        y = F.conv2d(tf32_round(x), tf32_round(w), b, padding=pad)
    elif kind == "bf16x3":
        x1, x2 = split_bf16(x, 2); w1, w2 = split_bf16(w, 2)
        y = F.conv2d(x1, w1, b, padding=pad) + F.conv2d(x2, w1, None, padding=pad) + F.conv2d(x1, w2, None, padding=pad)
    elif kind == "bf16x2":      # activations split, weights single bf16: x1*w1 + x2*w1
        x1, x2 = split_bf16(x, 2); (w1,) = split_bf16(w, 1)
        y = F.conv2d(x1, w1, b, padding=pad) + F.conv2d(x2, w1, None, padding=pad)
    elif kind == "f16x1":       # single fp16 operands (11-bit mantissa both)
        (x1,) = split_f16(x, 1); (w1,) = split_f16(w, 1)
        y = F.conv2d(x1, w1, b, padding=pad)
    elif kind == "f16x2":       # activations split in two fp16, weights single fp16: x1*w1 + x2*w1
        x1, x2 = split_f16(x, 2); (w1,) = split_f16(w, 1)
        y = F.conv2d(x1, w1, b, padding=pad) + F.conv2d(x2, w1, None, padding=pad)
    elif kind == "f16x2w":      # weights split, activations single fp16: x1*w1 + x1*w2
        (x1,) = split_f16(x, 1); w1, w2 = split_f16(w, 2)
        y = F.conv2d(x1, w1, b, padding=pad) + F.conv2d(x1, w2, None, padding=pad)
    elif kind == "f16x3":
        x1, x2 = split_f16(x, 2); w1, w2 = split_f16(w, 2)
        y = F.conv2d(x1, w1, b, padding=pad) + F.conv2d(x2, w1, None, padding=pad) + F.conv2d(x1, w2, None, padding=pad)
    elif kind == "bf16x6":
        x1, x2, x3 = split_bf16(x, 3); w1, w2, w3 = split_bf16(w, 3)
        y = F.conv2d(x1, w1, b, padding=pad)
        for xa, wa in ((x2, w1), (x1, w2), (x3, w1), (x2, w2), (x1, w3)):
            y = y + F.conv2d(xa, wa, None, padding=pad)
    else:
        raise ValueError(kind)
    return F.relu_(y) if relu else y


def run(label):
    torch.backends.cudnn.allow_tf32 = False
    cfgs = dict(superpoint_state=weights.make_superpoint_state(1), precision="bf16", conv_precision="f32")
    m = SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": 512, "match_threshold": 0.2,
                          "force_cpu": False, "sinkhorn_iterations": 20, "superglue_state": weights.make_superglue_state(2), **cfgs})
    m.match(g["image0"], g["image1"], quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
            geometric_verification=GeometricVerification.NONE)
    a = pairs(m.mkpts0, m.mkpts1)
    l = LightGlueMatcher({"features": "superpoint", "lightglue_state": weights.make_lightglue_state(3), **cfgs})
    l.match(g["image0"], g["image1"], quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[2, 2], overlap=40,
            max_keypoints=512, geometric_verification=GeometricVerification.NONE)
    b = pairs(l.mkpts0, l.mkpts1)
    print(f"{label:34s} SG match IoU {len(a & ref) / len(a | ref):.4f} (n={len(a)})   LG match IoU {len(b & refl) / len(b | refl):.4f} (n={len(b)})",
          flush=True)


orig_backbone = sp_mod.SuperPointB200.backbone
sp_mod.SuperPointB200._conv = emu_conv
sp_mod.SuperPointB200._pool = lambda self, x: F.max_pool2d(x, 2, 2)
KINDS = tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else ("f32", "tf32", "bf16x2", "bf16x3", "bf16x6")
for kind in KINDS:
    MODE.update(kind=kind, f32_layers=())
    run(f"backbone {kind}")
if len(sys.argv) > 1:
    sys.exit(0)
for layers in (("conv1a", "conv1b"), ("conv1a", "conv1b", "conv2a", "conv2b"), ("convPa", "convPb"), ("convPa", "convPb", "convDa", "convDb"),
               ("conv4a", "conv4b", "convPa", "convPb", "convDa", "convDb")):
    MODE.update(kind="tf32", f32_layers=layers)
    run("tf32 + f32 " + ",".join(layers))
