"""Features / Feature — drop-in for icepy4d/core/features.py:73-632 (the container the matcher's output is
poured into after every epoch, main_dev.py:165-190).

The reference keeps one Python `Feature` object per keypoint in a dict {track_id: Feature}; filling it costs a Python-level
constructor call, four isinstance asserts and a dict insertion per keypoint (core/features.py:441-454) — tens of thousands per
epoch, which becomes the wall-clock bottleneck of the workflow once the GPU path takes 85 ms.  Here the container is a
structure of arrays (x/y [n] f32, track ids [n] i32, descriptors [n, D] f32, scores [n] f32): `append_features_from_numpy` is a
handful of vectorised copies (the arrays that come off the device are stored as they are), and `Feature` objects are
materialised lazily, only for the keypoints a caller actually indexes.  Same public methods, argument checks, return types and
quirks as the reference:
  * descriptors go in and out as [D, n] (D in {128, 256}), keypoints as [n, 2] f32, scores as [n] f32;
  * duplicate incoming track ids are logged and replaced by progressive ids (core/features.py:415-426);
  * `append_features_from_numpy` with all-zero x returns without doing anything (`if not np.any(x)`, :385-387);
  * iteration (`__iter__` / `__next__`) looks features up by POSITION used as track id, like the reference (:288-299): it
    raises KeyError once a track id is missing (after a filter), exactly as the reference's dict lookup does;
  * filter_feature_by_index keeps the listed TRACK IDS (:574-582); filter_feature_by_mask is positional (:555-572).
"""
from __future__ import annotations

import logging
import pickle
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

logger = logging.getLogger(__name__)


def float32_type_check(array: np.ndarray, cast_integers: bool = False, verbose: bool = False) -> np.ndarray:
    """core/features.py:38-70: float64 (and, on request, int) arrays are cast to float32, anything else must already be f32."""
    if array.dtype == np.float64:
        if verbose:
            logger.info("Input array are float64 numbers. Casting them to np.float32")
        array = array.astype(np.float32)
    if cast_integers and array.dtype in (np.int32, np.int64):
        if verbose:
            logger.info("Input array are int numbers. Casting them to np.float32")
        array = array.astype(np.float32)
    if array.dtype != np.float32:
        raise ValueError("Invalid type of input array. It must be a numpy array of type np.float32")
    return array


class Feature:
    """One keypoint (core/features.py:73-205): x, y, track_id, descr ([D, 1] f32 or None), score, epoch — read-only properties."""

    __slots__ = ("_x", "_y", "_track", "_descr", "_score", "epoch")

    def __init__(self, x, y, track_id=None, descr: Optional[np.ndarray] = None, score=None, epoch=None) -> None:
        self._x = np.float32(x)
        self._y = np.float32(y)
        if track_id is not None:
            if isinstance(track_id, (int, np.int64)):
                track_id = np.int32(track_id)
            assert isinstance(track_id, np.int32), "Invalid track_id. It must be a integer number"
        self._track = track_id
        if descr is not None:
            msg = "Invalid descriptor. It must be a numpy array with lenght of 128 or 256."
            assert isinstance(descr, np.ndarray), msg
            assert descr.size in (128, 256) and descr.ndim <= 2, msg
            descr = float32_type_check(descr.reshape(-1, 1))                  # stored as a [D, 1] column, like the reference
        self._descr = descr
        self._score = None if score is None else np.float32(score)
        self.epoch = None if epoch is None else np.int32(epoch)

    def __repr__(self) -> str:
        return f"Feature with track_id={self._track}"

    @property
    def x(self) -> np.float32:
        return self._x

    @property
    def y(self) -> np.float32:
        return self._y

    @property
    def xy(self) -> np.ndarray:
        return np.array([self._x, self._y], dtype=np.float32).reshape(1, 2)

    @property
    def track_id(self):
        return self._track

    @property
    def descr(self):
        return self._descr

    @property
    def score(self):
        return self._score


class Features:
    """Collection of keypoints of one image (core/features.py:208-632), array-backed."""

    def __init__(self):
        self._descriptor_size = 256
        self.epoch = None
        self.reset_fetures()

    # ---- storage ------------------------------------------------------------------------------------------------
    def reset_fetures(self):                      # (sic) the reference's spelling
        self._xy = np.empty((0, 2), np.float32)
        self._ids = np.empty((0,), np.int32)
        self._descr = None                        # [n, D] f32, rows of NaN where a feature has no descriptor
        self._score = None                        # [n] f32, NaN where missing
        self._epoch = np.empty((0,), np.int64)    # -1 where missing
        self._last_id = -1
        self._iter = 0
        self._index: Optional[Dict[int, int]] = None

    def _rows(self) -> Dict[int, int]:
        if self._index is None:
            self._index = {int(t): i for i, t in enumerate(self._ids)}
        return self._index

    def _keep(self, rows: np.ndarray) -> None:
        self._xy, self._ids, self._epoch = self._xy[rows], self._ids[rows], self._epoch[rows]
        if self._descr is not None:
            self._descr = self._descr[rows]
        if self._score is not None:
            self._score = self._score[rows]
        self._index = None

    def _make(self, i: int) -> Feature:
        d = None if self._descr is None or np.isnan(self._descr[i, 0]) else self._descr[i]
        s = None if self._score is None or np.isnan(self._score[i]) else self._score[i]
        e = None if self._epoch[i] < 0 else self._epoch[i]
        return Feature(self._xy[i, 0], self._xy[i, 1], track_id=np.int32(self._ids[i]), descr=d, score=s, epoch=e)

    # ---- dict-like protocol of the reference -----------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._ids)

    def __getitem__(self, track_id) -> Optional[Feature]:
        i = self._rows().get(int(track_id))
        if i is None:
            logger.warning(f"Feature with track id {track_id} not available.")
            return None
        return self._make(i)

    def __contains__(self, track_id) -> bool:
        return int(track_id) in self._rows()

    def __delitem__(self, track_id) -> bool:
        i = self._rows().get(int(track_id))
        if i is None:
            logger.warning(f"Feature with track_id {track_id} not present")
            return False
        self._keep(np.delete(np.arange(len(self)), i))
        return True

    def __iter__(self):
        self._iter = 0
        return self

    def __next__(self) -> Feature:
        if self._iter < len(self):
            i = self._rows()[self._iter]          # KeyError if that track id is gone, like the reference's self._values[self._iter]
            self._iter += 1
            return self._make(i)
        self._iter = 0
        raise StopIteration

    def __repr__(self) -> str:
        return f"Features with {len(self)} features"

    @property
    def num_features(self) -> int:
        return len(self)

    @property
    def last_track_id(self):
        return self._last_id

    def get_track_ids(self) -> Tuple[np.int32, ...]:
        return tuple(np.int32(t) for t in self._ids)

    def set_last_track_id(self, last_track_id) -> None:
        try:
            self._last_id = np.int32(last_track_id)
        except Exception:
            raise ValueError("Invalid input argument last_track_id. It must be an integer number.")

    # ---- fill -----------------------------------------------------------------------------------------------------
    def _append(self, xy, ids, descr, scores, epoch) -> None:
        n0, n = len(self), len(ids)
        self._xy = np.concatenate([self._xy, xy])
        self._ids = np.concatenate([self._ids, np.asarray(ids, np.int32)])
        self._epoch = np.concatenate([self._epoch, np.full(n, -1 if epoch is None else int(epoch), np.int64)])
        if descr is not None or self._descr is not None:
            D = self._descriptor_size
            old = self._descr if self._descr is not None else np.full((n0, D), np.nan, np.float32)
            new = descr if descr is not None else np.full((n, D), np.nan, np.float32)
            self._descr = np.concatenate([old, new]) if n0 else np.ascontiguousarray(new)
        if scores is not None or self._score is not None:
            old = self._score if self._score is not None else np.full(n0, np.nan, np.float32)
            new = scores if scores is not None else np.full(n, np.nan, np.float32)
            self._score = np.concatenate([old, new])
        if self._index is not None:
            self._index.update({int(t): n0 + i for i, t in enumerate(ids)})
        if n:
            self._last_id = int(ids[-1])

    def append_feature(self, new_feature: Feature) -> None:
        """core/features.py:327-345: the feature is stored under the next progressive id."""
        assert isinstance(new_feature, Feature), "Invalid input feature. It must be Feature object"
        d = new_feature.descr
        if d is not None:
            if len(self) > 0 and self._descr is not None:
                assert self._descriptor_size == d.size, \
                    "Descriptor size of the new feature does not match with that of the existing feature"
            else:
                self._descriptor_size = d.size
        s = new_feature.score
        self._append(np.array([[new_feature.x, new_feature.y]], np.float32), [int(self._last_id) + 1],
                     None if d is None else d.reshape(1, -1), None if s is None else np.array([s], np.float32), new_feature.epoch)

    def append_features_from_numpy(self, x: np.ndarray, y: np.ndarray, descr: np.ndarray = None, scores: np.ndarray = None,
                                   track_ids: List[np.int32] = None, epoch: np.int32 = None) -> None:
        """core/features.py:362-454 — x, y [n] or [n,1]; descr [D, n]; scores [n] or [n,1]; track_ids: list of n ints."""
        assert isinstance(x, np.ndarray), "invalid type of x vector"
        assert isinstance(y, np.ndarray), "invalid type of y vector"
        if not np.any(x):
            logger.warning("Empty input feature arrays. Nothing done.")
            return None
        xx = float32_type_check(x, cast_integers=True).reshape(-1)
        yy = float32_type_check(y, cast_integers=True).reshape(-1)
        n = len(xx)
        if descr is not None:
            assert descr.shape[0] in (128, 256), \
                "invalid shape of the descriptor array. It must be of size mxn (m: descriptor size [128, 256], n: number of features"
            if len(self) > 0:
                assert self._descriptor_size == descr.shape[0], \
                    "Descriptor size of the new feature does not match with that of the existing feature"
            else:
                self._descriptor_size = descr.shape[0]
            descr = np.ascontiguousarray(float32_type_check(descr.T))
            assert descr.shape[0] >= n
            descr = descr[:n]
        if track_ids is None:
            ids = np.arange(int(self._last_id) + 1, int(self._last_id) + n + 1)
        else:
            assert isinstance(track_ids, list), \
                "Invalid track_ids input. It must be a list of integers of the same size of the input arrays."
            assert len(track_ids) == n, "invalid size of track_id input. It must be a list of the same size of the input arrays."
            ids = np.asarray(track_ids, dtype=np.int64)
            if len(self) and np.isin(ids, self._ids).any():
                dup = int(ids[np.isin(ids, self._ids)][0])
                logger.error(f"Feature with track_id {dup} is already present in Features object. Ignoring input track_id and "
                             "assigning progressive track_ids.")
                ids = np.arange(int(self._last_id) + 1, int(self._last_id) + n + 1)
        if scores is not None:
            scores = float32_type_check(scores).reshape(-1)[:n]
        if epoch is not None:
            try:
                epoch = np.int32(epoch)
            except Exception:
                raise ValueError("Invalid input argument epoch. It must be an integer number.")
            self.epoch = epoch
        self._append(np.stack([xx, yy], 1), ids, descr, scores, epoch)

    # ---- read-out -------------------------------------------------------------------------------------------------
    def kpts_to_numpy(self) -> np.ndarray:
        return self._xy.copy()

    def descr_to_numpy(self) -> np.ndarray:
        assert self._descr is not None and not np.isnan(self._descr[:, 0]).all(), "Descriptors non availble"
        return np.ascontiguousarray(self._descr.T)

    def scores_to_numpy(self) -> np.ndarray:
        assert self._score is not None and not np.isnan(self._score).all(), "Scores non availble"
        return self._score.copy()

    def to_numpy(self, get_descr: bool = False, get_score: bool = False) -> dict:
        """core/features.py:456-484 (scores are returned only together with the descriptors, as there)."""
        out = {"kpts": self.kpts_to_numpy()}
        if get_descr:
            out["descr"] = self.descr_to_numpy()
            if get_score:
                out["scores"] = self.scores_to_numpy()
        return out

    def get_features_as_dict(self, get_track_id: bool = False) -> dict:
        out = {"keypoints0": self.kpts_to_numpy(), "descriptors0": self.descr_to_numpy(), "scores0": self.scores_to_numpy()}
        if get_track_id:
            out["track_id"] = self.get_track_ids()
        return out

    # ---- filters --------------------------------------------------------------------------------------------------
    def filter_feature_by_mask(self, inlier_mask: List[bool]) -> None:
        m = np.array(inlier_mask)
        msg = ("It must be a boolean vector with the same lenght as the number of features stored in the Features object.")
        assert np.array_equal(m, m.astype(bool)), "Invalid type of input argument for inlier_mask. " + msg
        assert len(m) == len(self), "Invalid shape of input argument for inlier_mask. " + msg
        self._keep(np.flatnonzero(m.astype(bool)))

    def filter_feature_by_index(self, indexes: List[np.int32]) -> None:
        self._keep(np.flatnonzero(np.isin(self._ids, np.asarray(list(indexes), dtype=np.int64))))

    def get_feature_by_index(self, indexes: List[np.int32]) -> dict:
        keep = np.flatnonzero(np.isin(self._ids, np.asarray(list(indexes), dtype=np.int64)))
        return {int(self._ids[i]): self._make(i) for i in keep}

    # ---- on-disk formats ------------------------------------------------------------------------------------------
    def save_as_txt(self, path: Union[str, Path], fmt: str = "%i", delimiter: str = ",", header: str = "x,y"):
        np.savetxt(path, self.kpts_to_numpy(), fmt=fmt, delimiter=delimiter, newline="\n", header=header)

    def save_as_pickle(self, path: Union[str, Path]) -> bool:
        with open(Path(path), "wb") as f:
            pickle.dump(self, f, protocol=pickle.HIGHEST_PROTOCOL)
        return True
