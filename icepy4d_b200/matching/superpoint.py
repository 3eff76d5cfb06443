"""SuperPoint on the device: backbone + hand-written sm_100a post-processing.

Backbone precision policies (`conv_precision`):
  "f16x3" (default)  hand-written tcgen05 implicit-GEMM convolutions on split IEEE-half operands (csrc/conv_tc.cu): every f32 value
           travels as a (hi, lo) pair of halves (22 mantissa bits; values beyond the half range, +-65504, saturate) and three
           products are accumulated in f32: f32-grade results (~2^-22) at 16-bit tensor-core rate, 2x2 max-pools fused into
           the epilogues; end-to-end match IoU vs the reference 1.0000 / 0.9993 (scripts/precision_iou.py);
  "bf16x3"           the same kernels on split bf16 operands (16 mantissa bits, full f32 exponent range): 0.9979 / 0.9965;
  "f32"              cuDNN f32 (the on-device cross-check);   "tf32" / "f16" / "bf16"  cuDNN tensor-core convolutions.

Mirrors thirdparty/SuperGlue/models/superpoint.py:100-220 and thirdparty/LightGlue/lightglue/superpoint.py:88-215
of the reference; accepts a state_dict with the reference's parameter names.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .. import ops

_CONVS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "convPa", "convPb",
          "convDa", "convDb"]


@dataclass
class DeviceFeatures:
    """Keypoints of one image, resident in HBM.  Only the first `n` rows are valid."""
    keypoints: torch.Tensor      # [cap, 2] f32 (x, y)
    scores: torch.Tensor         # [cap] f32
    descriptors: torch.Tensor    # [cap, 256] f32, token-major (reference layout [256, n] is the transpose)
    n_dev: torch.Tensor          # int32[1] on the device
    n: Optional[int] = None      # host copy, filled by `sync_counts`
    height: int = 0
    width: int = 0


class SuperPointB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", nms_radius: int = 4,
                 keypoint_threshold: float = 0.005, max_keypoints: int = -1, remove_borders: int = 4,
                 conv_precision: str = "f16x3"):
        if not torch.cuda.is_available():
            raise RuntimeError("icepy4d_b200 needs a CUDA device (there is no CPU fallback)")
        assert conv_precision in ("f32", "tf32", "f16", "bf16", "bf16x3", "f16x3")
        self.device = torch.device(device)
        self.nms_radius, self.thr = int(nms_radius), float(keypoint_threshold)
        self.k = -1 if max_keypoints is None else int(max_keypoints)
        if self.k == 0 or self.k < -1:
            raise ValueError('"max_keypoints" must be positive or "-1"')
        self.border = int(remove_borders)
        self.conv_precision = conv_precision
        # "f16": fp16 operands / f32 accumulation — the same 10-bit operand mantissa as TF32 (the precision torch's cuDNN
        # convolutions use by default on the reference's CUDA path); the two 1x1 head convolutions stay f32/TF32 so that
        # logits and raw descriptors are produced in f32.
        self.act_dtype = {"bf16": torch.bfloat16, "f16": torch.float16}.get(conv_precision, torch.float32)
        self.w = {}
        self.pk = {}
        self.split = conv_precision in ("bf16x3", "f16x3")
        self.split_dtype = torch.float16 if conv_precision == "f16x3" else torch.bfloat16
        if self.split:
            for name in _CONVS[1:]:
                self.pk[name] = ops.PackedConv(state_dict[f"{name}.weight"], state_dict[f"{name}.bias"], self.device, self.split_dtype)
        for name in (_CONVS if not self.split else []):
            wdt = torch.float32 if name in ("convPb", "convDb") else self.act_dtype
            w = state_dict[f"{name}.weight"].to(self.device, dtype=wdt).contiguous(memory_format=torch.channels_last)
            b = state_dict[f"{name}.bias"].to(self.device, dtype=wdt)
            self.w[name] = (w, b)
        self.w1a = (state_dict["conv1a.weight"].to(self.device, torch.float32).contiguous(),
                    state_dict["conv1a.bias"].to(self.device, torch.float32).contiguous())
        self._kws = {}
        self.fuse_conv1 = os.environ.get("I4D_NO_CONV1_FUSION", "0") != "1"     # cross-check switch: separate conv1a / conv1b kernels

    # -- backbone (cuDNN through torch; channels-last so the heads come out HWC for the gather kernel) --
    def _conv(self, x, name, pad, relu=True):
        w, b = self.w[name]
        if relu and self.conv_precision != "f32":
            # cuDNN's fused conv+bias+ReLU epilogue: removes one full read+write of the activation per layer
            return torch.cudnn_convolution_relu(x, w, b, (1, 1), (pad, pad), (1, 1), 1)
        y = F.conv2d(x, w, b, padding=pad)
        return F.relu_(y) if relu else y

    def _pool(self, x):
        if self.conv_precision != "f32" and x.is_contiguous(memory_format=torch.channels_last):
            return ops.maxpool2x2_cl(x)           # hand-written: one 16-byte load per input pixel-channel-quad
        return F.max_pool2d(x, 2, 2)

    def _backbone_split(self, image: torch.Tensor):
        pk = self.pk
        if self.fuse_conv1:
            # conv1a computed inside conv1b's operand producer: its 1 GB (2000 x 2000 tile) activation never goes to HBM
            x = ops.sp_conv1ab_fused(image, self.w1a[0], self.w1a[1], pk["conv1b"], pool=True)
        else:
            x = ops.sp_conv1a_relu_split(image, self.w1a[0], self.w1a[1], self.split_dtype)
            x = ops.conv_bf16x3(x, pk["conv1b"], pool=True)
        x = ops.conv_bf16x3(ops.conv_bf16x3(x, pk["conv2a"]), pk["conv2b"], pool=True)
        x = ops.conv_bf16x3(ops.conv_bf16x3(x, pk["conv3a"]), pk["conv3b"], pool=True)
        x = ops.conv_bf16x3(ops.conv_bf16x3(x, pk["conv4a"]), pk["conv4b"])
        logits = ops.conv_bf16x3(ops.conv_bf16x3(x, pk["convPa"]), pk["convPb"], relu=False, out="planar")
        desc = ops.conv_bf16x3(ops.conv_bf16x3(x, pk["convDa"]), pk["convDb"], relu=False, out="nhwc")
        return logits, desc

    def backbone(self, image: torch.Tensor):
        """image [1,1,H,W] f32 -> (logits [65,h,w] f32 contiguous, desc [h,w,256] f32 contiguous)."""
        if self.split:
            return self._backbone_split(image.contiguous())
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = self.conv_precision != "f32"
        try:
            if self.conv_precision == "f32":
                x = image.contiguous(memory_format=torch.channels_last)
                x = self._conv(x, "conv1a", 1)
            else:
                # hand-written first layer: exact f32 FMAs, written straight in channels-last (no 1 GB transpose)
                x = ops.sp_conv1a_relu(image, self.w1a[0], self.w1a[1], self.act_dtype)
            x = self._conv(x, "conv1b", 1)
            x = self._pool(x)
            x = self._conv(self._conv(x, "conv2a", 1), "conv2b", 1)
            x = self._pool(x)
            x = self._conv(self._conv(x, "conv3a", 1), "conv3b", 1)
            x = self._pool(x)
            x = self._conv(self._conv(x, "conv4a", 1), "conv4b", 1)
            logits = self._conv(self._conv(x, "convPa", 1).float(), "convPb", 0, relu=False)
            desc = self._conv(self._conv(x, "convDa", 1).float(), "convDb", 0, relu=False)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        logits = logits[0].contiguous()                              # [65,h,w] planar
        desc = desc[0].permute(1, 2, 0).contiguous()                 # [h,w,256] (no copy when channels-last)
        return logits, desc

    def postprocess(self, logits: torch.Tensor, desc_hwc: torch.Tensor, k: Optional[int] = None) -> DeviceFeatures:
        k = self.k if k is None else int(k)
        scores = ops.sp_score_map(logits)
        H, W = scores.shape
        key = (H, W)
        if key not in self._kws:
            self._kws[key] = ops.KeypointWorkspace(H, W, self.device)
        kpts, ksc, n_dev, _ = ops.sp_keypoints(scores, self.nms_radius, self.thr, self.border, k, self._kws[key])
        if k < 0:
            # "keep all": the result size is data dependent -> one host read to trim the buffers
            n = min(int(n_dev.item()), kpts.shape[0])
            kpts, ksc = kpts[:n].contiguous(), ksc[:n].contiguous()
            return DeviceFeatures(kpts, ksc, ops.sp_sample_descriptors(desc_hwc, kpts, None, n), n_dev, n)
        desc = ops.sp_sample_descriptors(desc_hwc, kpts, n_dev)
        return DeviceFeatures(kpts, ksc, desc, n_dev)

    def detect(self, image: torch.Tensor, k: Optional[int] = None) -> DeviceFeatures:
        """image [1,1,H,W] f32 in [0,1] on the device."""
        logits, desc = self.backbone(image)
        f = self.postprocess(logits, desc, k)
        f.height, f.width = int(image.shape[-2]), int(image.shape[-1])
        return f


def sync_counts(*feats: DeviceFeatures) -> None:
    """One device->host read for the keypoint counts of several images."""
    if not feats:
        return
    counts = torch.cat([f.n_dev for f in feats]).cpu().tolist()
    for f, c in zip(feats, counts):
        f.n = min(int(c), f.keypoints.shape[0])
