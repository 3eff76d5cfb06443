"""TEST INFRASTRUCTURE — CPU oracle (torch fp32) for the SuperPoint stages.  Never imported by the product.

Restates, stage by stage, what the reference computes (citations are into /root/reference/src/icepy4d):
  backbone            thirdparty/SuperGlue/models/superpoint.py:154-168,206-207  (LG copy: LightGlue/lightglue/superpoint.py:155-169,203-204)
  score_map           superpoint.py:169-172   softmax over 65 channels, drop dustbin, 8x8 pixel shuffle
  simple_nms          superpoint.py:48-64     3-pass max-pool NMS
  keypoints_sg        superpoint.py:176-203   threshold -> border removal -> top-k -> (x, y)
  keypoints_lg        LightGlue/lightglue/superpoint.py:176-200  borders set to -1 *before* the threshold
  sample_descriptors  superpoint.py:82-97,208 dense L2 norm, bilinear grid_sample(align_corners=True), L2 norm
Pinned against the reference itself by tests/test_oracle_vs_golden.py (fixtures from oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def backbone(image: torch.Tensor, sd: dict):
    """image [1,1,H,W] f32 in [0,1] -> (logits [65,h,w], raw descriptors [256,h,w])."""
    def cr(x, name, pad):
        return F.relu(F.conv2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=pad))

    x = cr(cr(image, "conv1a", 1), "conv1b", 1)
    x = F.max_pool2d(x, 2, 2)
    x = cr(cr(x, "conv2a", 1), "conv2b", 1)
    x = F.max_pool2d(x, 2, 2)
    x = cr(cr(x, "conv3a", 1), "conv3b", 1)
    x = F.max_pool2d(x, 2, 2)
    x = cr(cr(x, "conv4a", 1), "conv4b", 1)
    logits = F.conv2d(cr(x, "convPa", 1), sd["convPb.weight"], sd["convPb.bias"])
    desc = F.conv2d(cr(x, "convDa", 1), sd["convDb.weight"], sd["convDb.bias"])
    return logits[0], desc[0]


def score_map(logits: torch.Tensor) -> torch.Tensor:
    """[65,h,w] -> [8h,8w]; channel c lands at (dy, dx) = (c // 8, c % 8) inside its 8x8 cell."""
    p = torch.softmax(logits, 0)[:64]
    _, h, w = p.shape
    return p.reshape(8, 8, h, w).permute(2, 0, 3, 1).reshape(8 * h, 8 * w)


def simple_nms(s: torch.Tensor, r: int) -> torch.Tensor:
    def mp(x):
        return F.max_pool2d(x[None, None], 2 * r + 1, 1, r)[0, 0]

    keep = s == mp(s)
    for _ in range(2):
        supp = mp(keep.float()) > 0
        s2 = torch.where(supp, torch.zeros_like(s), s)
        keep = keep | ((s2 == mp(s2)) & ~supp)
    return torch.where(keep, s, torch.zeros_like(s))


def keypoints_sg(nms: torch.Tensor, thr: float, border: int, k: int):
    """Returns (kpts [N,2] f32 as (x,y), scores [N]).  k < 0 disables top-k."""
    H, W = nms.shape
    yx = torch.nonzero(nms > thr)
    sc = nms[yx[:, 0], yx[:, 1]]
    ok = (yx[:, 0] >= border) & (yx[:, 0] < H - border) & (yx[:, 1] >= border) & (yx[:, 1] < W - border)
    yx, sc = yx[ok], sc[ok]
    if 0 <= k < len(sc):
        sc, idx = torch.topk(sc, k)
        yx = yx[idx]
    return yx.flip(1).float(), sc


def keypoints_lg(nms: torch.Tensor, thr: float, border: int, k):
    s = nms.clone()
    if border:
        s[:border] = -1
        s[:, :border] = -1
        s[-border:] = -1
        s[:, -border:] = -1
    ys, xs = torch.where(s > thr)
    sc = s[ys, xs]
    yx = torch.stack([ys, xs], -1)
    if k is not None and k < len(sc):
        sc, idx = torch.topk(sc, k, sorted=True)
        yx = yx[idx]
    return yx.flip(1).float(), sc


def sample_descriptors(kpts: torch.Tensor, desc_raw: torch.Tensor) -> torch.Tensor:
    """kpts [N,2] (x,y); desc_raw [256,h,w] -> [256,N] unit-norm columns."""
    c, h, w = desc_raw.shape
    d = F.normalize(desc_raw[None], p=2, dim=1)
    g = (kpts - 3.5) / torch.tensor([w * 8 - 4.5, h * 8 - 4.5]) * 2 - 1
    out = F.grid_sample(d, g.view(1, 1, -1, 2), mode="bilinear", align_corners=True)
    return F.normalize(out.reshape(1, c, -1), p=2, dim=1)[0]


def superpoint_sg(image, sd, nms_radius=3, thr=0.001, k=-1, border=4):
    """Whole SuperGlue-flavour SuperPoint: returns dict like the reference forward() minus the batch lists."""
    logits, desc = backbone(image, sd)
    nms = simple_nms(score_map(logits), nms_radius)
    kp, sc = keypoints_sg(nms, thr, border, k)
    return {"keypoints": kp, "scores": sc, "descriptors": sample_descriptors(kp, desc),
            "logits": logits, "desc_raw": desc}


def superpoint_lg(image, sd, nms_radius=4, thr=0.0005, k=None, border=4):
    """LightGlue-flavour SuperPoint.forward (image already grey, [1,1,H,W]); descriptors [N,256]."""
    logits, desc = backbone(image, sd)
    nms = simple_nms(score_map(logits), nms_radius)
    kp, sc = keypoints_lg(nms, thr, border, k)
    return {"keypoints": kp, "keypoint_scores": sc, "descriptors": sample_descriptors(kp, desc).t().contiguous(),
            "logits": logits, "desc_raw": desc}
