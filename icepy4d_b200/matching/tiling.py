"""Tile grid arithmetic — same limits as the reference's Tiler (icepy4d/matching/tiling.py:93-135), including its
quirks: Python banker's `round`, tile index = row * ncol + col, and end-exclusive slicing of the inclusive-looking
(xmin, ymin, xmax, ymax) limits, so a tile is DX + overlap - 1 pixels wide."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


class Tiler:
    def __init__(self, grid: List[int] = [1, 1], overlap: int = 0, origin: List[int] = [0, 0], max_length: int = 2000) -> None:
        self._origin = origin
        self._overlap = overlap
        self._nrow, self._ncol = grid[0], grid[1]
        self._limits = None

    @property
    def grid(self) -> List[int]:
        return [self._nrow, self._ncol]

    @property
    def origin(self) -> List[int]:
        return self._origin

    @property
    def overlap(self) -> int:
        return self._overlap

    @property
    def limits(self) -> Dict[int, tuple]:
        return self._limits

    def compute_limits_by_shape(self, h: int, w: int) -> Tuple[Dict[int, tuple], List[int]]:
        ox, oy = self._origin[0], self._origin[1]
        dx = round((w - ox) / self._ncol / 10) * 10
        dy = round((h - oy) / self._nrow / 10) * 10
        lims = {}
        for col in range(self._ncol):
            for row in range(self._nrow):
                xmin = max(ox, col * dx - self._overlap)
                ymin = max(oy, row * dy - self._overlap)
                lims[row * self._ncol + col] = (xmin, ymin, xmin + dx + self._overlap - 1, ymin + dy + self._overlap - 1)
        self._limits = lims
        return lims, self._origin

    def compute_limits_by_grid(self, image: np.ndarray):
        return self.compute_limits_by_shape(image.shape[0], image.shape[1])

    @staticmethod
    def patch_shape(limits, h: int, w: int) -> Tuple[int, int, int, int]:
        """(x0, y0, tw, th) of the end-exclusive slice image[ymin:ymax, xmin:xmax], clipped like numpy slicing."""
        x0, y0 = limits[0], limits[1]
        x1, y1 = min(limits[2], w), min(limits[3], h)
        return x0, y0, max(0, x1 - x0), max(0, y1 - y0)

    def extract_patch(self, image: np.ndarray, limits) -> np.ndarray:
        return image[limits[1]:limits[3], limits[0]:limits[2]]
