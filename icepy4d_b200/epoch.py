"""Per-epoch stereo pipeline (match -> verify -> triangulate) and epoch sharding across GPUs.

One epoch = one stereo pair through the reference's per-epoch calls (main_dev.py:115-132, 220-251 of the reference):
`matcher.match(...)` (tiling, SuperPoint, SuperGlue/LightGlue, merge, geometric verification) followed by
`Triangulate(...).triangulate_two_views()`.  Epochs share no state on this path, so multi-GPU execution is a static
partition of epoch indices over ranks (one process per GPU) with no collective on the data path and a single
result gather at the end.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .matching import GeometricVerification, LightGlueMatcher, Quality, SuperGlueMatcher, TileSelection


@dataclass
class EpochResult:
    mkpts0: np.ndarray
    mkpts1: np.ndarray
    scores0: np.ndarray
    points3d: np.ndarray
    status: np.ndarray
    F: Optional[np.ndarray]


class StereoEpochPipeline:
    def __init__(self, matcher, cameras: Sequence, quality: Quality = Quality.HIGH,
                 tile_selection: TileSelection = TileSelection.GRID, **match_config):
        self.matcher = matcher
        self.cameras = list(cameras)
        self.quality, self.tile_selection = quality, tile_selection
        self.match_config = dict(match_config)
        self.match_config.setdefault("geometric_verification", GeometricVerification.MAGSAC)

    # -- device-resident: inputs already in HBM, results stay in HBM (kernel-only throughput) --
    def run_device(self, dev0: torch.Tensor, dev1: torch.Tensor) -> Dict[str, torch.Tensor]:
        if self.quality != Quality.HIGH:                     # matchers.py:583-610 on the device (bit-exact pyrDown / pyrUp)
            dev0, dev1 = self.matcher._resize_images_device(self.quality, dev0, dev1)
        mk0, mk1, s0, s1, conf, d0, d1, F = self.matcher.match_device(dev0, dev1, self.quality, self.tile_selection,
                                                                      **self.match_config)
        out = {"mkpts0": mk0, "mkpts1": mk1, "scores0": s0, "F": F}
        if mk0.shape[0] > 0:
            u0 = ops.undistort_points(mk0.contiguous(), self.cameras[0].K, self.cameras[0].dist)
            u1 = ops.undistort_points(mk1.contiguous(), self.cameras[1].K, self.cameras[1].dist)
            out["points3d"], out["status"] = ops.triangulate_iterative_ls(u0, u1, self.cameras[0].P, self.cameras[1].P)
        else:
            out["points3d"] = torch.zeros((0, 3), dtype=torch.float64, device=mk0.device)
            out["status"] = torch.zeros((0,), dtype=torch.int32, device=mk0.device)
        return out

    # -- end to end through the public plugin API: host images in, host arrays out --
    def run(self, image0: np.ndarray, image1: np.ndarray) -> EpochResult:
        from .sfm import Triangulate

        m = self.matcher
        m.match(image0, image1, quality=self.quality, tile_selection=self.tile_selection, **self.match_config)
        tri = Triangulate(self.cameras, [m.mkpts0, m.mkpts1])
        X = tri.triangulate_two_views()
        return EpochResult(m.mkpts0, m.mkpts1, m.scores0, X, tri.status, m._F)


def shard_epochs(n_epochs: int, rank: int, world_size: int) -> List[int]:
    """Static round-robin partition: rank r owns {e : e mod world_size == r}."""
    return list(range(rank, n_epochs, world_size))


def gather_results(local: Dict[int, np.ndarray], world_size: int, group=None) -> Dict[int, np.ndarray]:
    """The single end-of-run exchange: every rank contributes {epoch: array}; all ranks get the union.
    Works with gloo (CPU) and nccl (object gather goes through pinned host staging)."""
    import torch.distributed as dist

    if world_size == 1 or not dist.is_initialized():
        return dict(local)
    bucket: List[Optional[dict]] = [None] * world_size
    dist.all_gather_object(bucket, local, group=group)
    out: Dict[int, np.ndarray] = {}
    for part in bucket:
        out.update(part)
    return out


def gather_results_device(local: Dict[int, torch.Tensor], world_size: int, group=None) -> Dict[int, torch.Tensor]:
    """The same single end-of-run exchange for device-resident results, without pickling: {epoch: [n_e, C] tensor} per rank ->
    the union on every rank.  Two collectives in total: the (epoch id, row count) tables, then ONE all_gather of the rank's
    results concatenated and padded to the longest rank (NCCL over NVLink on GPUs; gloo on CPU tensors).  Ranks may hold
    different numbers of epochs and rows; C and dtype must agree."""
    import torch.distributed as dist

    if world_size == 1 or not dist.is_initialized():
        return dict(local)
    keys = sorted(local)
    ref = local[keys[0]] if keys else None
    dev = ref.device if ref is not None else torch.device("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    # table exchange: [n_epochs, (epoch, rows)...] padded to the largest table
    n_max = torch.tensor([len(keys)], dtype=torch.int64, device=dev)
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=group)
    n_max = int(n_max.item())
    meta = torch.full((1 + 2 * n_max + 2,), -1, dtype=torch.int64, device=dev)
    meta[0] = len(keys)
    for i, k in enumerate(keys):
        meta[1 + 2 * i], meta[2 + 2 * i] = int(k), int(local[k].shape[0])
    cols = int(np.prod(ref.shape[1:])) if ref is not None else 1
    meta[-2], meta[-1] = cols, sum(int(local[k].shape[0]) for k in keys)
    metas = [torch.empty_like(meta) for _ in range(world_size)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu() for m in metas]
    rows_max = max(int(m[-1]) for m in metas)
    cols = max(int(m[-2]) for m in metas)
    dtype = ref.dtype if ref is not None else torch.float64
    flat = torch.zeros((max(rows_max, 1), cols), dtype=dtype, device=dev)
    if keys:
        cat = torch.cat([local[k].reshape(local[k].shape[0], cols) for k in keys])
        flat[: cat.shape[0]] = cat
    bucket = [torch.empty_like(flat) for _ in range(world_size)]
    dist.all_gather(bucket, flat, group=group)
    out: Dict[int, torch.Tensor] = {}
    for m, buf in zip(metas, bucket):
        off = 0
        for i in range(int(m[0])):
            e, r = int(m[1 + 2 * i]), int(m[2 + 2 * i])
            out[e] = buf[off:off + r]
            off += r
    return out


def make_cfg1_pipeline(max_keypoints: int = 2048, precision: str = "f32", conv_precision: str = "f16x3", cameras=None) -> StereoEpochPipeline:
    """BASELINE.json configs[0]: the notebook's single stereo epoch — SuperPoint + LightGlue, 2048 kp, 6000x4000 downsampled to
    1500x1000 (Quality.LOW = two pyrDown), one tile (GRID [1,1]; TileSelection.NONE is buggy in the reference, App. D.1)."""
    from . import synthetic, weights

    m = LightGlueMatcher({"features": "superpoint", "superpoint_state": weights.make_superpoint_state(1),
                          "lightglue_state": weights.make_lightglue_state(3), "precision": precision,
                          "conv_precision": conv_precision})
    if cameras is None:
        cameras = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    return StereoEpochPipeline(m, cameras, Quality.LOW, TileSelection.GRID, grid=[1, 1], overlap=0,
                               max_keypoints=max_keypoints, geometric_verification=GeometricVerification.MAGSAC)


def make_cfg2_pipeline(max_keypoints: int = 8192, sinkhorn_iterations: int = 100, precision: str = "f32",
                       conv_precision: str = "f16x3", grid=(2, 3), cameras=None) -> StereoEpochPipeline:
    """BASELINE.json configs[1]: full-res 6000x4000 stereo pair, 2x3 tile grid, SuperPoint + SuperGlue (outdoor
    architecture, 100 Sinkhorn iterations), 8192 keypoints per tile, seeded structured random weights."""
    from . import synthetic, weights

    m = SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": max_keypoints,
                          "match_threshold": 0.2, "force_cpu": False, "sinkhorn_iterations": sinkhorn_iterations,
                          "superpoint_state": weights.make_superpoint_state(1), "superglue_state": weights.make_superglue_state(2),
                          "precision": precision, "conv_precision": conv_precision})
    if cameras is None:
        cameras = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    return StereoEpochPipeline(m, cameras, Quality.HIGH, TileSelection.GRID, grid=list(grid), overlap=0,
                               geometric_verification=GeometricVerification.MAGSAC)


def make_cfg5_pipeline(max_keypoints: int = 16384, precision: str = "f32", conv_precision: str = "f16x3", grid=(3, 4),
                       cameras=None) -> StereoEpochPipeline:
    """BASELINE.json configs[4]: 16384 kp/tile LightGlue, 3x4 tiles, dual-softmax + mutual-NN (static depth/width)."""
    from . import synthetic, weights

    m = LightGlueMatcher({"features": "superpoint", "superpoint_state": weights.make_superpoint_state(1),
                          "lightglue_state": weights.make_lightglue_state(3), "precision": precision,
                          "conv_precision": conv_precision, "depth_confidence": -1, "width_confidence": -1})
    if cameras is None:
        cameras = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    return StereoEpochPipeline(m, cameras, Quality.HIGH, TileSelection.GRID, grid=list(grid), overlap=0,
                               max_keypoints=max_keypoints, geometric_verification=GeometricVerification.MAGSAC)
