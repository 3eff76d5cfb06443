"""Matcher plugin surface — drop-in for icepy4d.matching.{SuperGlueMatcher, LightGlueMatcher} (reference:
icepy4d/matching/matchers.py:51-700, 826-940, 1202-1342): same constructor options, `match()` signature / kwargs,
result properties and error conventions.  Underneath, both images are uploaded once, every tile goes
u8 -> grey f32 -> SuperPoint -> matcher on the device, and only the merged, verified matches come back.

Differences that are deliberate (documented in DESIGN.md):
  * there is no CPU path: `force_cpu=True` raises instead of silently running something else;
  * tiles are detected once and cached when a tile takes part in several pairs (same results, less work);
  * weights come from `opt["superpoint_state"]` / `opt["superglue_state"]` / `opt["lightglue_state"]`
    (state_dict or path) or from $ICEPY4D_WEIGHTS_DIR — the reference's bundled .pth files are not redistributable;
  * descriptors of the matches stay in HBM until `.descriptors0/1` is read.
"""
from __future__ import annotations

import logging
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from itertools import product
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import cv2
import numpy as np
import torch

from .. import ops
from .enums import GeometricVerification, Quality, TileSelection
from .geometric_verification import geometric_verification_device
from .superpoint import DeviceFeatures, SuperPointB200, sync_counts
from .tiling import Tiler

logger = logging.getLogger(__name__)

MIN_MATCHES_PER_TILE = 5


def check_dict_keys(dict: dict, keys: List[str]):
    missing_keys = [key for key in keys if key not in dict]
    if missing_keys:
        raise KeyError(f"Missing required keys: {', '.join(missing_keys)} Matcher option dictionary")


@dataclass
class FeaturesBase:
    keypoints: np.ndarray
    descriptors: np.ndarray = None
    scores: np.ndarray = None


def _load_state(spec, default_name: str):
    if isinstance(spec, dict):
        return spec
    path = spec
    if path is None:
        wdir = os.environ.get("ICEPY4D_WEIGHTS_DIR")
        if wdir:
            path = os.path.join(wdir, default_name)
    if path is None or not os.path.exists(str(path)):
        raise FileNotFoundError(
            f"weights '{default_name}' not found: pass a state_dict/path in the matcher options or set ICEPY4D_WEIGHTS_DIR")
    return torch.load(str(path), map_location="cpu")


class ImageMatcherABC(ABC):
    @abstractmethod
    def match(self):
        pass

    @abstractmethod
    def _match_images(self):
        pass

    @abstractmethod
    def _match_by_tile(self):
        pass


@dataclass
class _PairResult:
    """Per tile pair, fixed-size device arrays (no host sync): row i is valid iff valid[i]."""
    mk0: torch.Tensor
    mk1: torch.Tensor
    s0: torch.Tensor
    s1: torch.Tensor
    conf: torch.Tensor
    valid: torch.Tensor
    d0: torch.Tensor      # [n0,256] descriptors of image-0 keypoints
    d1: torch.Tensor      # [n1,256]
    m0: torch.Tensor      # [n0] int32


class ImageMatcherBase(ImageMatcherABC):
    GRAY_MODE = 0

    def __init__(self, opt: dict = {}) -> None:
        if not isinstance(opt, dict):
            raise TypeError("opt must be a dictionary")
        self._opt = dict(opt)
        if opt.get("force_cpu"):
            raise RuntimeError("icepy4d_b200 has no CPU path: force_cpu=True is not supported")
        if not torch.cuda.is_available():
            raise RuntimeError("icepy4d_b200 needs a CUDA device (there is no CPU fallback)")
        self._device = "cuda"
        logger.info(f"Running inference on device {self._device}")
        self.reset()
        self._F = None
        self._do_viz = False
        self._save_dir = None

    def reset(self):
        self._mkpts0 = self._mkpts1 = None
        self._descriptors0 = self._descriptors1 = None
        self._scores0 = self._scores1 = None
        self._mconf = None
        self._desc_dev = None

    # -- result properties (matchers.py:107-137) --
    @property
    def device(self):
        return self._device

    @property
    def mkpts0(self):
        return self._mkpts0

    @property
    def mkpts1(self):
        return self._mkpts1

    def _materialise_descriptors(self):
        if self._desc_dev is not None:
            d0, d1 = self._desc_dev
            self._descriptors0 = d0.t().contiguous().cpu().numpy()   # [256,S] like the reference
            self._descriptors1 = d1.t().contiguous().cpu().numpy()
            self._desc_dev = None

    @property
    def descriptors0(self):
        self._materialise_descriptors()
        return self._descriptors0

    @property
    def descriptors1(self):
        self._materialise_descriptors()
        return self._descriptors1

    @property
    def scores0(self):
        return self._scores0

    @property
    def scores1(self):
        return self._scores1

    @property
    def mconf(self):
        return self._mconf

    # -- main entry point (matchers.py:140-261) --
    def match(self, image0: np.ndarray, image1: np.ndarray, quality: Quality = Quality.HIGH,
              tile_selection: TileSelection = TileSelection.NONE, **config) -> bool:
        gv_method = config.get("geometric_verification", GeometricVerification.PYDEGENSAC)
        threshold = config.get("threshold", 1)
        confidence = config.get("confidence", 0.9999)
        self._do_viz = config.get("do_viz_matches", False)
        save_dir = config.get("save_dir", None)
        self._save_dir = Path(save_dir) if save_dir is not None else None
        if self._save_dir is not None:
            self._save_dir.mkdir(parents=True, exist_ok=True)
        if self._do_viz:
            logger.warning("visualisation is outside the B200 hot path; do_viz_matches is ignored")

        if quality == Quality.HIGH and tile_selection in (TileSelection.GRID, TileSelection.EXHAUSTIVE):
            # tiles are cut straight from the uploaded images: upload in row bands on a copy stream (band order = tile order) and
            # let every tile wait only for its own rows, so SuperPoint starts after the first band instead of after 144 MB
            dev0, dev1 = self._upload_banded(image0, image1, config.get("grid", [1, 1]))
        else:
            dev0, dev1 = self._resize_images_device(quality, self._upload(image0), self._upload(image1))
        try:
            mk0, mk1, s0, s1, conf, d0, d1, F = self.match_device(dev0, dev1, quality, tile_selection, **config)
        finally:
            self._row_events = {}
        self._F = F.cpu().numpy().reshape(3, 3) if F is not None else None
        if self._F is not None and not np.isfinite(self._F).all():    # no model found: the mask was all ones (degrade path)
            self._F = None
        self._store_device_results(mk0, mk1, s0, s1, conf, d0, d1)
        if self._save_dir is not None:
            self.save_mkpts_as_txt(self._save_dir)
        return True

    def match_device(self, dev0: torch.Tensor, dev1: torch.Tensor, quality: Quality = Quality.HIGH,
                     tile_selection: TileSelection = TileSelection.NONE, **config):
        """Device-resident core of `match()`: u8 images already in HBM (already resized for `quality`) ->
        (mkpts0, mkpts1, scores0, scores1, mconf, desc0 [S,256], desc1 [S,256], F [9] f64 or None), all on the device."""
        gv_method = config.get("geometric_verification", GeometricVerification.PYDEGENSAC)
        threshold = config.get("threshold", 1)
        confidence = config.get("confidence", 0.9999)
        if tile_selection == TileSelection.NONE:
            logger.info("Matching full images...")
            res = self._match_tensors(dev0, dev1, (0, 0, dev0.shape[1], dev0.shape[0]), (0, 0, dev1.shape[1], dev1.shape[0]), **config)
            mk0, mk1, s0, s1, conf, d0, d1 = self._merge_pairs([res], [[0, 0]], [[0, 0]], dedupe=False)
            merged = (mk0, mk1, s0, s1, self._full_image_mconf(s0, conf), d0, d1)
        else:
            logger.info("Matching by tiles...")
            merged = self._match_by_tile_device(dev0, dev1, tile_selection, **config)
        mk0, mk1, s0, s1, conf, d0, d1 = merged

        scale = {Quality.HIGHEST: 0.5, Quality.HIGH: 1.0, Quality.MEDIUM: 2.0, Quality.LOW: 4.0}[quality]
        if scale != 1.0:   # _resize_features (matchers.py:612-639)
            mk0, mk1 = mk0 * scale, mk1 * scale
        logger.info("Matching done!")

        F = None
        if gv_method is not GeometricVerification.NONE:
            logger.info("Performing geometric verification...")
            F, mask = self._verify(mk0, mk1, gv_method, threshold, confidence)
            idx = torch.nonzero(mask).squeeze(1)
            mk0, mk1, s0, s1, conf, d0, d1 = (t.index_select(0, idx) for t in (mk0, mk1, s0, s1, conf, d0, d1))
            logger.info("Geometric verification done.")
        return mk0, mk1, s0, s1, conf, d0, d1, F

    def _full_image_mconf(self, s0, conf):
        return s0.clone()          # SuperGlue: mconf = features0.scores[valid] (matchers.py:937-938)

    def _verify(self, mk0, mk1, gv_method, threshold, confidence):
        if mk0.shape[0] < 4:   # geometric_verification.py:50-52
            logger.warning("Not enough matches to perform geometric verification.")
            return None, torch.ones(mk0.shape[0], dtype=torch.bool, device=mk0.device)
        try:
            return geometric_verification_device(mk0, mk1, gv_method, threshold, confidence)
        except Exception as err:
            logger.error(f"{err}. Unable to perform geometric verification.")
            return None, torch.ones(mk0.shape[0], dtype=torch.bool, device=mk0.device)

    def _store_device_results(self, mk0, mk1, s0, s1, conf, d0, d1):
        pack = torch.cat([mk0, mk1, s0[:, None], s1[:, None], conf[:, None]], 1).cpu().numpy()   # one D2H copy
        self._mkpts0 = np.ascontiguousarray(pack[:, 0:2])
        self._mkpts1 = np.ascontiguousarray(pack[:, 2:4])
        self._scores0 = np.ascontiguousarray(pack[:, 4])
        self._scores1 = np.ascontiguousarray(pack[:, 5])
        self._mconf = np.ascontiguousarray(pack[:, 6])
        self._descriptors0 = self._descriptors1 = None
        self._desc_dev = (d0, d1)

    # -- host-side helpers --
    def _upload(self, image: np.ndarray) -> torch.Tensor:
        assert isinstance(image, np.ndarray), "images must be NumPy arrays"
        if image.dtype != np.uint8:
            raise TypeError("images must be uint8 (H x W or H x W x 3)")
        if image.ndim == 3 and image.shape[2] not in (1, 3):
            raise ValueError(f"Not an image: {image.shape}")
        t = torch.from_numpy(np.ascontiguousarray(image))
        return t.cuda(non_blocking=True)

    def _upload_banded(self, image0: np.ndarray, image1: np.ndarray, grid):
        """H2D of both images in `grid[0]` row bands each, alternating between the images (the order in which the tile loop
        touches them), on a dedicated copy stream.  Every band records an event; `_tile_tensor` makes the compute stream wait
        for the bands a tile overlaps.  With pageable host memory the copies are synchronous and this degrades to `_upload`."""
        imgs = []
        for image in (image0, image1):
            assert isinstance(image, np.ndarray), "images must be NumPy arrays"
            if image.dtype != np.uint8:
                raise TypeError("images must be uint8 (H x W or H x W x 3)")
            if image.ndim == 3 and image.shape[2] not in (1, 3):
                raise ValueError(f"Not an image: {image.shape}")
            imgs.append(torch.from_numpy(np.ascontiguousarray(image)))
        devs = [torch.empty(t.shape, dtype=torch.uint8, device="cuda") for t in imgs]
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        cs, cur = self._copy_stream, torch.cuda.current_stream()
        cs.wait_stream(cur)                                           # the destination buffers belong to the compute stream
        nb = max(1, int(grid[0]))
        self._row_events = {d.data_ptr(): [] for d in devs}
        with torch.cuda.stream(cs):
            for b in range(nb):
                for t, d in zip(imgs, devs):
                    h = t.shape[0]
                    r0, r1 = (h * b) // nb, (h * (b + 1)) // nb
                    d[r0:r1].copy_(t[r0:r1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    self._row_events[d.data_ptr()].append((r1, ev))
        for d in devs:
            d.record_stream(cs)
        return devs[0], devs[1]

    def _wait_rows(self, dev_img: torch.Tensor, y1: int):
        """Compute stream waits for the upload bands covering rows [0, y1) of a banded image (no-op otherwise)."""
        evs = getattr(self, "_row_events", {}).get(dev_img.data_ptr())
        if not evs:
            return
        cur = torch.cuda.current_stream()
        while evs:
            r1, ev = evs[0]
            cur.wait_event(ev)
            evs.pop(0)
            if r1 >= y1:
                break

    def _resize_images_device(self, quality: Quality, dev0: torch.Tensor, dev1: torch.Tensor):
        """matchers.py:583-610 on the device: bit-exact cv2.pyrUp / pyrDown kernels on the uploaded u8 images."""
        if quality == Quality.HIGHEST:
            return ops.pyr_up(dev0), ops.pyr_up(dev1)
        if quality == Quality.MEDIUM:
            return ops.pyr_down(dev0), ops.pyr_down(dev1)
        if quality == Quality.LOW:
            return ops.pyr_down(ops.pyr_down(dev0)), ops.pyr_down(ops.pyr_down(dev1))
        return dev0, dev1

    def _resize_images(self, quality: Quality, image0: np.ndarray, image1: np.ndarray):
        # host variant kept for API compatibility (matchers.py:583-610); match() uses the device kernels above.
        if quality == Quality.HIGHEST:
            return cv2.pyrUp(image0), cv2.pyrUp(image1)
        if quality == Quality.MEDIUM:
            return cv2.pyrDown(image0), cv2.pyrDown(image1)
        if quality == Quality.LOW:
            return cv2.pyrDown(cv2.pyrDown(image0)), cv2.pyrDown(cv2.pyrDown(image1))
        return image0, image1

    def _tile_tensor(self, dev_img: torch.Tensor, rect) -> torch.Tensor:
        x0, y0, tw, th = rect
        self._wait_rows(dev_img, y0 + th)
        return ops.tile_to_gray_f32(dev_img, x0, y0, tw, th, self.GRAY_MODE)

    # -- subclass hooks --
    def _detect(self, tile: torch.Tensor, **config) -> DeviceFeatures:
        raise NotImplementedError

    def _match_features(self, f0: DeviceFeatures, f1: DeviceFeatures, **config):
        """-> (matches0 [n0] i32, conf [n0] f32) on the device"""
        raise NotImplementedError

    def _match_tensors(self, dev0, dev1, rect0, rect1, **config) -> _PairResult:
        f0 = self._detect(self._tile_tensor(dev0, rect0), **config)
        f1 = self._detect(self._tile_tensor(dev1, rect1), **config)
        sync_counts(f0, f1)
        return self._pair(f0, f1, **config)

    def _pair(self, f0: DeviceFeatures, f1: DeviceFeatures, **config) -> _PairResult:
        n0, n1 = f0.n, f1.n
        k0, k1 = f0.keypoints[:n0], f1.keypoints[:n1]
        m0, conf = self._match_features(f0, f1, **config)
        valid = m0 > -1
        j = m0.clamp(min=0).long()
        if n1 == 0:
            z = torch.zeros((n0, 2), device=k0.device)
            return _PairResult(k0, z, f0.scores[:n0], torch.zeros(n0, device=k0.device), conf, valid, f0.descriptors[:n0], f1.descriptors[:n1], m0)
        return _PairResult(k0, k1.index_select(0, j), f0.scores[:n0], f1.scores[:n1].index_select(0, j), conf, valid,
                           f0.descriptors[:n0], f1.descriptors[:n1], m0)

    def _merge_pairs(self, results: List[_PairResult], offs0, offs1, dedupe: bool, keep_all: bool = False):
        """Concatenate per-pair results, add tile offsets, drop unmatched rows, optionally de-duplicate on exact
        image-0 coordinates keeping the first occurrence and sort lexicographically by (x, y) — what
        np.unique(axis=0, return_index=True) does in the reference (matchers.py:441-450)."""
        dev = torch.device("cuda")
        if not results:
            e2, e1, ed = torch.zeros((0, 2), device=dev), torch.zeros(0, device=dev), torch.zeros((0, 256), device=dev)
            return e2, e2.clone(), e1, e1.clone(), e1.clone(), ed, ed.clone()
        # tile offsets are added as scalars to row ranges of the concatenated copies: a `torch.tensor(offset, device=...)` per pair is a
        # blocking pageable host->device copy each (12 of them per cfg2 epoch, 20-40 us of idle GPU apiece in the merge tail)
        def cat_with_offsets(parts, offs):
            out = torch.cat(parts)
            a = 0
            for t, o in zip(parts, offs):
                b = a + t.shape[0]
                for c in (0, 1):
                    if o[c] != 0 and b > a:
                        out[a:b, c] += float(o[c])
                a = b
            return out

        mk0 = cat_with_offsets([r.mk0 for r in results], offs0)
        mk1 = cat_with_offsets([r.mk1 for r in results], offs1)
        s0, s1 = torch.cat([r.s0 for r in results]), torch.cat([r.s1 for r in results])
        conf = torch.cat([r.conf for r in results])
        valid = torch.cat([r.valid for r in results])
        d0 = torch.cat([r.d0 for r in results])
        base1 = np.cumsum([0] + [r.d1.shape[0] for r in results[:-1]])
        j1 = torch.cat([r.m0.clamp(min=0).long() + int(b) for r, b in zip(results, base1)])
        d1_all = torch.cat([r.d1 for r in results])
        if keep_all:
            idx = torch.arange(mk0.shape[0], device=dev)
        else:
            idx = torch.nonzero(valid).squeeze(1)                # the one data-dependent size of the epoch
        if dedupe and idx.numel() > 0:
            sel = mk0.index_select(0, idx)
            key = sel[:, 0].to(torch.int64) * (1 << 24) + sel[:, 1].to(torch.int64)
            skey, order = torch.sort(key, stable=True)
            first = torch.ones_like(skey, dtype=torch.bool)
            first[1:] = skey[1:] != skey[:-1]
            idx = idx.index_select(0, order[first])
        pick = lambda t: t.index_select(0, idx)
        d1 = d1_all.index_select(0, j1.index_select(0, idx)) if d1_all.shape[0] else torch.zeros((idx.numel(), 256), device=dev)
        return pick(mk0), pick(mk1), pick(s0), pick(s1), pick(conf), pick(d0), d1

    # -- tiling (matchers.py:304-469) --
    def _match_by_tile_device(self, dev0, dev1, tile_selection, **config):
        grid = config.get("grid", [1, 1])
        overlap = config.get("overlap", 0)
        origin = config.get("origin", [0, 0])
        self._tiler = Tiler(grid=grid, overlap=overlap, origin=origin)
        h0, w0, h1, w1 = dev0.shape[0], dev0.shape[1], dev1.shape[0], dev1.shape[1]
        t0_lims, t0_origin = self._tiler.compute_limits_by_shape(h0, w0)
        t1_lims, t1_origin = self._tiler.compute_limits_by_shape(h1, w1)
        tile_pairs = self._tile_selection(dev0, dev1, t0_lims, t1_lims, tile_selection, config=config)
        feats0: Dict[int, DeviceFeatures] = {}
        feats1: Dict[int, DeviceFeatures] = {}
        for t0, t1 in tile_pairs:                                 # SuperPoint on every tile that takes part, once
            if t0 not in feats0:
                feats0[t0] = self._detect(self._tile_tensor(dev0, Tiler.patch_shape(t0_lims[t0], h0, w0)), **config)
            if t1 not in feats1:
                feats1[t1] = self._detect(self._tile_tensor(dev1, Tiler.patch_shape(t1_lims[t1], h1, w1)), **config)
        sync_counts(*feats0.values(), *feats1.values())           # one D2H read for all keypoint counts
        results, offs0, offs1 = [], [], []
        for t0, t1 in tile_pairs:
            logger.info(f" - Matching tile pair ({t0}, {t1})")
            results.append(self._pair(feats0[t0], feats1[t1], **config))
            offs0.append([t0_lims[t0][0] + t0_origin[0], t0_lims[t0][1] + t0_origin[1]])
            offs1.append([t1_lims[t1][0] + t1_origin[0], t1_lims[t1][1] + t1_origin[1]])
        mk0, mk1, s0, s1, conf, d0, d1 = self._merge_pairs(results, offs0, offs1, dedupe=True)
        return mk0, mk1, s0, s1, s0.clone(), d0, d1              # tiled mconf = features0.scores (matchers.py:464-465)

    def _tile_selection(self, dev0, dev1, t0_lims, t1_lims, method: TileSelection = TileSelection.PRESELECTION, **config):
        # NOTE: the reference passes its kwargs as ONE kwarg named `config` (matchers.py:353-355), so
        # `min_matches_per_tile` always falls back to the default; kept (Appendix D.2).
        min_matches_per_tile = config.get("min_matches_per_tile", MIN_MATCHES_PER_TILE)
        if method == TileSelection.EXHAUSTIVE:
            return sorted(product(t0_lims.keys(), t1_lims.keys()))
        if method == TileSelection.GRID:
            return sorted(zip(t0_lims.keys(), t1_lims.keys()))
        if method == TileSelection.PRESELECTION:
            h = dev0.shape[0]
            n_down = 3 if h > 4000 else (2 if h > 2000 else 1)     # matchers.py:516-523 (the >8000 branch is dead code there)
            i0, i1 = dev0, dev1
            for _ in range(n_down):
                i0, i1 = ops.pyr_down(i0), ops.pyr_down(i1)           # on the device, bit-exact with cv2.pyrDown
            r = self._match_tensors(i0, i1, (0, 0, i0.shape[1], i0.shape[0]), (0, 0, i1.shape[1], i1.shape[0]), max_keypoints=4096)
            # the decision itself (strict rectangle test :495-499, count > min_matches :555) on the device: one small kernel
            # over the pre-matches, then T0 x T1 integers come back — no keypoint array crosses to the host
            k0, k1 = sorted(t0_lims.keys()), sorted(t1_lims.keys())
            counts = ops.tile_pair_counts(r.mk0.contiguous(), r.mk1.contiguous(), r.valid, float(2 ** n_down),
                                          [t0_lims[k] for k in k0], [t1_lims[k] for k in k1]).cpu().numpy()
            self._preselection_counts = counts
            return [(a, b) for ia, a in enumerate(k0) for ib, b in enumerate(k1) if counts[ia, ib] > min_matches_per_tile]
        raise ValueError(f"unsupported tile selection {method}")

    # -- numpy-level API kept for compatibility (matchers.py:276-302, 304-469) --
    def _match_images(self, image0: np.ndarray, image1: np.ndarray, **config):
        d0, d1 = self._upload(image0), self._upload(image1)
        r = self._match_tensors(d0, d1, (0, 0, d0.shape[1], d0.shape[0]), (0, 0, d1.shape[1], d1.shape[0]), **config)
        return self._pair_to_numpy(r)

    def _pair_to_numpy(self, r: _PairResult):
        m0 = r.m0.cpu().numpy().astype(np.int64)
        k1 = r.d1.shape[0]
        f0 = FeaturesBase(r.mk0.cpu().numpy(), r.d0.t().contiguous().cpu().numpy(), r.s0.cpu().numpy())
        f1 = FeaturesBase(self._last_f1_kpts.cpu().numpy(), r.d1.t().contiguous().cpu().numpy(), self._last_f1_scores.cpu().numpy())
        return f0, f1, m0, self._mconf_from(r, m0)

    def _mconf_from(self, r: _PairResult, m0: np.ndarray):
        return r.s0.cpu().numpy()[m0 > -1]

    def _match_by_tile(self, image0: np.ndarray, image1: np.ndarray, tile_selection: TileSelection = TileSelection.PRESELECTION, **config):
        assert isinstance(image0, np.ndarray), "image0 must be a NumPy array"
        assert isinstance(image1, np.ndarray), "image1 must be a NumPy array"
        mk0, mk1, s0, s1, conf, d0, d1 = self._match_by_tile_device(self._upload(image0), self._upload(image1), tile_selection, **config)
        f0 = FeaturesBase(mk0.cpu().numpy(), d0.t().contiguous().cpu().numpy(), s0.cpu().numpy())
        f1 = FeaturesBase(mk1.cpu().numpy(), d1.t().contiguous().cpu().numpy(), s1.cpu().numpy())
        return f0, f1, np.arange(mk0.shape[0]), f0.scores.copy()

    def save_mkpts_as_txt(self, savedir: Union[str, Path], delimiter: str = ",", header: str = "x,y") -> None:
        path = Path(savedir)
        path.mkdir(parents=True, exist_ok=True)
        np.savetxt(path / "keypoints_0.txt", self.mkpts0, delimiter=delimiter, newline="\n", header=header)
        np.savetxt(path / "keypoints_1.txt", self.mkpts1, delimiter=delimiter, newline="\n", header=header)


class SuperGlueMatcher(ImageMatcherBase):
    """Drop-in for icepy4d.matching.SuperGlueMatcher (matchers.py:826-940)."""
    GRAY_MODE = 0

    def __init__(self, opt: dict) -> None:
        cfg = self._build_superglue_config(opt)
        super().__init__(opt)
        self._cfg = cfg
        sp_state = _load_state(opt.get("superpoint_state"), "superpoint_v1.pth")
        sg_state = _load_state(opt.get("superglue_state"), f"superglue_{cfg['superglue']['weights']}.pth")
        self.superpoint = SuperPointB200(sp_state, self._device, nms_radius=cfg["superpoint"]["nms_radius"],
                                         keypoint_threshold=cfg["superpoint"]["keypoint_threshold"],
                                         max_keypoints=cfg["superpoint"]["max_keypoints"],
                                         conv_precision=opt.get("conv_precision", "f16x3"))
        from .superglue import SuperGlueB200
        self.superglue = SuperGlueB200(sg_state, self._device, sinkhorn_iterations=cfg["superglue"]["sinkhorn_iterations"],
                                       match_threshold=cfg["superglue"]["match_threshold"],
                                       precision=opt.get("precision", "f32"))

    def _build_superglue_config(self, opt: dict) -> dict:
        def_opt = {"weights": "outdoor", "keypoint_threshold": 0.001, "max_keypoints": -1, "match_threshold": 0.3,
                   "force_cpu": False, "nms_radius": 3, "sinkhorn_iterations": 20}
        opt = {**def_opt, **opt}
        check_dict_keys(opt, ["weights", "keypoint_threshold", "max_keypoints", "match_threshold", "force_cpu"])
        assert opt["weights"] in ["indoor", "outdoor"]
        return {"superpoint": {"nms_radius": opt["nms_radius"], "keypoint_threshold": opt["keypoint_threshold"],
                               "max_keypoints": opt["max_keypoints"]},
                "superglue": {"weights": opt["weights"], "sinkhorn_iterations": opt["sinkhorn_iterations"],
                              "match_threshold": opt["match_threshold"]},
                "force_cpu": opt["force_cpu"]}

    def _detect(self, tile, **config):
        return self.superpoint.detect(tile)     # SG ignores the per-call max_keypoints kwarg (Appendix D.3)

    def _match_features(self, f0, f1, **config):
        n0, n1 = f0.n, f1.n
        self._last_f1_kpts, self._last_f1_scores = f1.keypoints[:n1], f1.scores[:n1]
        m0, m1, ms0, ms1 = self.superglue.match(f0.keypoints[:n0], f0.scores[:n0], f0.descriptors[:n0], (f0.height, f0.width),
                                                f1.keypoints[:n1], f1.scores[:n1], f1.descriptors[:n1], (f1.height, f1.width))
        return m0, ms0


class LightGlueMatcher(ImageMatcherBase):
    """Drop-in for icepy4d.matching.LightGlueMatcher (matchers.py:1202-1342)."""
    GRAY_MODE = 1

    def __init__(self, opt: dict = {}) -> None:
        self._localfeatures = opt.get("features", "superpoint")
        if self._localfeatures != "superpoint":
            raise NotImplementedError("only SuperPoint features are on the B200 hot path (DISK is out of scope)")
        super().__init__(opt)
        sp_state = _load_state(opt.get("superpoint_state"), "superpoint_v1.pth")
        lg_state = _load_state(opt.get("lightglue_state"), "superpoint_lightglue.pth")
        self._sp_state = sp_state
        self._sp_cache: Dict[int, SuperPointB200] = {}
        from .lightglue import LightGlueB200
        self.lightglue = LightGlueB200(lg_state, self._device, precision=opt.get("precision", "f32"),
                                       depth_confidence=opt.get("depth_confidence", 0.95),
                                       width_confidence=opt.get("width_confidence", 0.99),
                                       filter_threshold=opt.get("filter_threshold", 0.1))

    def _full_image_mconf(self, s0, conf):
        # matches01["scores"] (matchers.py:1290).  NOTE: with TileSelection.NONE the reference's LightGlueMatcher
        # stores ALL keypoints of both images, unmatched and of different lengths (Appendix D.1); that bug is not
        # reproduced: only the matched pairs are stored here.
        return conf

    def _extractor(self, k: int) -> SuperPointB200:
        if k not in self._sp_cache:   # LG-flavour SuperPoint: radius 4, threshold 0.0005 (LightGlue/lightglue/superpoint.py:97-103)
            self._sp_cache[k] = SuperPointB200(self._sp_state, self._device, nms_radius=4, keypoint_threshold=0.0005,
                                               max_keypoints=k, conv_precision=self._opt.get("conv_precision", "f16x3"))
        return self._sp_cache[k]

    def _detect(self, tile, **config):
        if config.get("resize", None) is not None:
            raise NotImplementedError("resize != None (kornia antialiased resize) is outside the B200 hot path")
        return self._extractor(int(config.get("max_keypoints", 10240))).detect(tile)

    def _match_features(self, f0, f1, **config):
        n0, n1 = f0.n, f1.n
        self._last_f1_kpts, self._last_f1_scores = f1.keypoints[:n1], f1.scores[:n1]
        out = self.lightglue.match(f0.keypoints[:n0], f0.descriptors[:n0], (f0.width, f0.height),
                                   f1.keypoints[:n1], f1.descriptors[:n1], (f1.width, f1.height))
        self._last_lg = out
        return out["matches0"], out["matching_scores0"]

    def _mconf_from(self, r, m0):
        return self._last_lg["scores"].cpu().numpy()      # matches01["scores"] (matchers.py:1290)
