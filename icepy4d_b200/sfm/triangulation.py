"""Triangulate — drop-in for icepy4d/sfm/triangulation.py:43-116 (same constructor, method names, arguments and
return values).  Undistortion and triangulation run as one-thread-per-point f64 CUDA kernels; the cameras only need
the attributes the reference reads: `.K`, `.dist`, `.P` (core/camera.py:117-207)."""
from __future__ import annotations

import logging
from typing import List

import numpy as np
import torch

from .. import ops


class Triangulate:
    def __init__(self, cameras: List = None, image_points: List[np.ndarray] = None) -> None:
        self.cameras = cameras
        self.image_points = image_points
        self.points3d = None
        self.colors = None
        self.status = None

    def _undistorted(self, view: int) -> torch.Tensor:
        pts = torch.as_tensor(np.ascontiguousarray(self.image_points[view], dtype=np.float32)).cuda()
        cam = self.cameras[view]
        return ops.undistort_points(pts, cam.K, cam.dist)

    def triangulate_two_views(self, views_ids: List[int] = [0, 1], approach: str = "iterative_LS_triangulation",
                              compute_colors: bool = False, image: np.ndarray = None, cam_id: int = 0) -> np.ndarray:
        n = len(self.image_points[views_ids[0]])
        if n != len(self.image_points[views_ids[1]]):
            raise ValueError("Number of points don't match.")
        if n == 0:
            self.points3d = np.zeros((0, 3))
            return self.points3d
        u0, u1 = self._undistorted(views_ids[0]), self._undistorted(views_ids[1])
        P0, P1 = self.cameras[views_ids[0]].P, self.cameras[views_ids[1]].P
        if approach == "iterative_LS_triangulation":
            X, status = ops.triangulate_iterative_ls(u0, u1, P0, P1)
            self.status = status.cpu().numpy()
            logging.info(f"Point triangulation succeded: {(self.status == 1).sum() / max(1, self.status.size)}.")
            self.points3d = X.cpu().numpy()
            if compute_colors:
                assert image is not None and type(image) == np.ndarray, "Invalid input image for interpolating point colors"
                self.interpolate_colors_from_image(image, self.cameras[cam_id], _points_dev=X)
        elif approach == "linear_triangulation":
            self.points3d = ops.triangulate_dlt(u0, u1, P0, P1).cpu().numpy()
        return self.points3d

    def interpolate_colors_from_image(self, image: np.ndarray, camera, convert_BRG2RGB: bool = True, _points_dev=None):
        """sfm/triangulation.py:133-148 + sfm/interpolate_colors.py:14-51: project the triangulated points into `image` with the
        camera's Brown model and interpolate the colours bilinearly (values in [0, 1], Nx(channels) f64)."""
        assert self.points3d is not None, "points 3D are not available, Triangulate homologous points first."
        assert image.ndim == 3, "invalid input image. Image has not 3 channel"
        X = _points_dev if _points_dev is not None else torch.as_tensor(np.ascontiguousarray(self.points3d, dtype=np.float64)).cuda()
        img = torch.as_tensor(np.ascontiguousarray(image, dtype=np.uint8)).cuda()
        self.colors = ops.interpolate_point_colors(X.contiguous(), img, camera.R, camera.t, camera.K, camera.dist,
                                                   convert_bgr2rgb=convert_BRG2RGB).cpu().numpy()
        logging.info("Point colors interpolated")
        return self.colors
