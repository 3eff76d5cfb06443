"""Points / Point — drop-in for icepy4d/core/points.py:76-514: the triangulated 3-D points of an epoch with
their track ids and colours.  Array-backed like `Features` (coordinates [n,3] f32, colours [n,3] f32 in [0,1], ids [n] i32):
`append_points_from_numpy` stores what `Triangulate` returns with three vectorised copies instead of one Python object per point
(core/points.py:364-366).  Same methods, checks and quirks as the reference; `to_point_cloud` (Open3D) is out of scope."""
from __future__ import annotations

import logging
import pickle
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from .features import float32_type_check


class Point:
    """core/points.py:76-170: coordinates (3,) f32, track_id, color (3,) f32 in [0,1], cov."""

    def __init__(self, coordinates: np.ndarray, track_id: int = None, color: np.ndarray = None, cov: np.ndarray = None) -> None:
        assert isinstance(coordinates, np.ndarray), "invalid argument coordinates"
        if coordinates.shape in ((3, 1), (1, 3)):
            coordinates = coordinates.reshape(3)
        assert coordinates.shape == (3,), "Invalid shape of coordinates array. It must be a (3,) numpy array (vector)"
        coordinates = float32_type_check(coordinates, cast_integers=True)
        assert isinstance(color, np.ndarray), "invalid argument color"          # the reference requires a colour (:119)
        if color.shape in ((3, 1), (1, 3)):
            color = color.reshape(3)
        assert color.shape == (3,), "Invalid shape of color array. It must be a (3,) numpy array"
        color = float32_type_check(color, cast_integers=False)
        self._track_id, self._color, self._cov = track_id, color, cov
        self._X, self._Y, self._Z = coordinates

    track_id = property(lambda self: self._track_id)
    X = property(lambda self: self._X)
    Y = property(lambda self: self._Y)
    Z = property(lambda self: self._Z)
    color = property(lambda self: self._color)

    @property
    def coordinates(self) -> np.ndarray:
        return np.array([self._X, self._Y, self._Z], dtype=np.float32)

    def project(self, camera) -> np.ndarray:
        """core/points.py:159-169: pinhole projection with the camera's P (no distortion)."""
        x = np.asarray(camera.P, np.float64) @ np.append(self.coordinates.astype(np.float64), 1.0)
        return (x[:2] / x[2]).astype(np.float32)


class Points:
    def __init__(self):
        self.reset_points()

    def reset_points(self):
        self._xyz = np.empty((0, 3), np.float32)
        self._col = np.empty((0, 3), np.float32)      # NaN rows where a point has no colour
        self._ids = np.empty((0,), np.int32)
        self._last_id = -1
        self._iter = 0
        self._index: Optional[Dict[int, int]] = None

    def _rows(self) -> Dict[int, int]:
        if self._index is None:
            self._index = {int(t): i for i, t in enumerate(self._ids)}
        return self._index

    def _keep(self, rows: np.ndarray) -> None:
        self._xyz, self._col, self._ids = self._xyz[rows], self._col[rows], self._ids[rows]
        self._index = None

    def _make(self, i: int) -> Point:
        return Point(self._xyz[i], int(self._ids[i]), color=self._col[i])

    def __len__(self) -> int:
        return len(self._ids)

    def __getitem__(self, track_id) -> Optional[Point]:
        i = self._rows().get(int(track_id))
        if i is None:
            logging.warning(f"Point with track id {track_id} not available.")
            return None
        return self._make(i)

    def __contains__(self, track_id) -> bool:
        return int(track_id) in self._rows()

    def __delitem__(self, track_id) -> bool:
        i = self._rows().get(int(track_id))
        if i is None:
            logging.warning(f"Point with track_id {track_id} not present")
            return False
        self._keep(np.delete(np.arange(len(self)), i))
        return True

    def __iter__(self):
        self._iter = 0
        return self

    def __next__(self) -> Point:
        if self._iter < len(self):
            i = self._rows()[self._iter]
            self._iter += 1
            return self._make(i)
        self._iter = 0
        raise StopIteration

    def __repr__(self):
        return f"Points with {len(self)} points"

    @property
    def num_points(self) -> int:
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

    def append_point(self, new_point: Point) -> None:
        assert isinstance(new_point, Point), "Invalid input point. It must be Point object"
        self.append_points_from_numpy(new_point.coordinates.reshape(1, 3), colors=new_point.color.reshape(1, 3))

    def append_points_from_numpy(self, coordinates: np.ndarray, track_ids: List[np.int32] = None, colors: np.ndarray = None) -> None:
        """core/points.py:317-368 — coordinates [n,3]; track_ids list/tuple of n ints; colors [n,3] floats in [0,1]."""
        if not np.any(coordinates):
            logging.warning("Empty input feature arrays. Nothing done.")
            return None
        assert isinstance(coordinates, np.ndarray), "invalid argument coordinates"
        assert coordinates.shape[1] == 3, "Invalid shape of coordinates array. It must be a nx3 numpy array"
        coordinates = float32_type_check(coordinates, cast_integers=True)
        n = len(coordinates)
        if track_ids is None:
            ids = np.arange(int(self._last_id) + 1, int(self._last_id) + n + 1)
        else:
            assert isinstance(track_ids, (list, tuple)), \
                "Invalid track_ids input. It must be a list or a tuple of integers of the same size of the input arrays."
            assert len(track_ids) == n, "invalid size of track_id input. It must be a list of the same size of the input arrays."
            ids = np.asarray(track_ids, dtype=np.int64)
            if len(self) and np.isin(ids, self._ids).any():
                dup = int(ids[np.isin(ids, self._ids)][0])
                logging.error(f"Feature with track_id {dup} is already present in Features object. Ignoring input track_id and "
                              "assigning progressive track_ids.")
                ids = np.arange(int(self._last_id) + 1, int(self._last_id) + n + 1)
        col = np.full((n, 3), np.nan, np.float32) if colors is None else np.float32(colors).reshape(n, 3)
        n0 = len(self)
        self._xyz = np.concatenate([self._xyz, coordinates])
        self._col = np.concatenate([self._col, col])
        self._ids = np.concatenate([self._ids, ids.astype(np.int32)])
        if self._index is not None:
            self._index.update({int(t): n0 + i for i, t in enumerate(ids)})
        self._last_id = int(ids[-1])

    def to_numpy(self) -> np.ndarray:
        return self._xyz.copy()

    def colors_to_numpy(self, as_uint8: bool = False) -> np.ndarray:
        return np.uint8(self._col * 255) if as_uint8 else self._col.copy()

    def filter_point_by_mask(self, inlier_mask: List[bool], verbose: bool = False) -> None:
        m = np.asarray(inlier_mask)
        msg = "It must be a boolean vector with the same lenght as the number of points stored in the Points object."
        assert np.array_equal(m, m.astype(bool)), "Invalid type of input argument for inlier_mask. " + msg
        assert len(m) == len(self), "Invalid shape of input argument for inlier_mask. " + msg
        self.filter_points_by_index([int(t) for t in self._ids[m.astype(bool)]], verbose=verbose)

    def filter_points_by_index(self, indexes: List[np.int32], verbose: bool = False) -> None:
        keep = np.flatnonzero(np.isin(self._ids, np.asarray(list(indexes), dtype=np.int64)))
        if verbose:
            logging.info(f"Points filtered: {len(self) - len(keep)}/{len(self)} removed. New Points size: {len(keep)}.")
        last_id = int(self._ids[keep[-1]])             # IndexError on an empty selection, like the reference (:458)
        self._keep(keep)
        self._last_id = last_id

    def get_points_by_index(self, indexes: List[np.int32]) -> dict:
        keep = np.flatnonzero(np.isin(self._ids, np.asarray(list(indexes), dtype=np.int64)))
        return {int(self._ids[i]): self._make(i) for i in keep}

    def save_as_txt(self, path: Union[str, Path], fmt: str = "%i", delimiter: str = ",", header: str = "x,y"):        # (sic) the reference default
        np.savetxt(path, self.to_numpy(), fmt=fmt, delimiter=delimiter, newline="\n", header=header)

    def save_as_pickle(self, path: Union[str, Path]) -> bool:
        with open(Path(path), "wb") as f:
            pickle.dump(self, f, protocol=pickle.HIGHEST_PROTOCOL)
        return True
