"""Absolute_orientation — drop-in for icepy4d/sfm/absolute_orientation.py:56-287: Helmert (similarity)
transformation between the local reconstruction and world coordinates, and its application to points and cameras.

What runs where: the reductions over the point sets (`i4d_helmert_moments`) and the transformation of point clouds
(`i4d_apply_transform`) are sm_100a kernels (f64); the 4x4 symmetric eigenproblem of Horn's quaternion method and the Euler-angle
extraction are a handful of host flops, restated from the vendored thirdparty/transformations.py:889-1020, 1274-1310, 1040-1100.
`estimate_transformation_least_squares` needs lmfit, which the reference imports lazily and which is absent from its own
environment here: like the reference it raises ImportError when lmfit is missing, and it is outside the hot path otherwise.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

from .. import ops
from .triangulation import Triangulate

_EPS = np.finfo(float).eps * 4.0


def quaternion_matrix(q: np.ndarray) -> np.ndarray:
    """thirdparty/transformations.py:1274-1310 (w, x, y, z) -> homogeneous rotation matrix."""
    q = np.array(q, dtype=np.float64, copy=True)
    n = float(np.dot(q, q))
    if n < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([[1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0], 0.0],
                     [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0], 0.0],
                     [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2], 0.0],
                     [0.0, 0.0, 0.0, 1.0]])


def euler_from_matrix_sxyz(M: np.ndarray) -> Tuple[float, float, float]:
    """thirdparty/transformations.py euler_from_matrix(matrix, axes="sxyz")."""
    M = np.asarray(M, dtype=np.float64)[:3, :3]
    cy = math.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > _EPS:
        return math.atan2(M[2, 1], M[2, 2]), math.atan2(-M[2, 0], cy), math.atan2(M[1, 0], M[0, 0])
    return math.atan2(-M[1, 2], M[1, 1]), math.atan2(-M[2, 0], cy), 0.0


def similarity_from_moments(mom: np.ndarray, estimate_scale: bool = True) -> np.ndarray:
    """The closed form of affine_matrix_from_points(v0, v1, shear=False, scale=estimate_scale, usesvd=False) from the 17 moments
    of `ops.helmert_moments` (transformations.py:989-1020): quaternion = eigenvector of the largest eigenvalue of Horn's N."""
    t0, t1 = -mom[0:3], -mom[3:6]
    C = mom[6:15].reshape(3, 3)                        # C[a, b] = sum c0[a] c1[b]
    xx, yy, zz = C[0, 0], C[1, 1], C[2, 2]
    xy, yz, zx = C[0, 1], C[1, 2], C[2, 0]             # v0 * roll(v1, -1): (x0 y1, y0 z1, z0 x1)
    xz, yx, zy = C[0, 2], C[1, 0], C[2, 1]             # v0 * roll(v1, -2): (x0 z1, y0 x1, z0 y1)
    N = np.array([[xx + yy + zz, 0.0, 0.0, 0.0],
                  [yz - zy, xx - yy - zz, 0.0, 0.0],
                  [zx - xz, xy + yx, yy - xx - zz, 0.0],
                  [xy - yx, zx + xz, yz + zy, zz - xx - yy]])
    w, V = np.linalg.eigh(N)                           # lower triangle, like the reference
    q = V[:, np.argmax(w)]
    q = q / np.linalg.norm(q)
    M = quaternion_matrix(q)
    if estimate_scale:
        M[:3, :3] *= math.sqrt(mom[16] / mom[15])
    M0, M1 = np.identity(4), np.identity(4)
    M0[:3, 3], M1[:3, 3] = t0, t1
    M = np.linalg.inv(M1) @ (M @ M0)
    return M / M[3, 3]


class Absolute_orientation:
    def __init__(self, cameras: Tuple, points3d_final: np.ndarray, points3d_orig: np.ndarray = None,
                 image_points: Tuple[np.ndarray] = None, camera_centers_world: Tuple[np.ndarray] = None) -> None:
        self.cameras = cameras
        if points3d_final is not None and points3d_final.shape[1] == 3:
            self.v1 = points3d_final
        else:
            raise ValueError("Missing or wrong input for points in world reference system. Please, provide the their 3D "
                             "coordinates in nx3 numpy array.")
        if points3d_orig is not None:
            self.v0 = points3d_orig
        elif image_points is not None:
            self.v0 = self.triangulate_image_points(image_points)
        else:
            raise ValueError("Missing input for points in local reference system. Please, provide the their 3D coordinates or "
                             "the image points to be triangulated")
        if camera_centers_world is not None:
            self.add_camera_centers_to_points(camera_centers_world)

    def add_camera_centers_to_points(self, camera_centers_world: List, v0: np.ndarray = None, v1: np.ndarray = None) -> None:
        # absolute_orientation.py:100-137 (incl. its swapped assignment of the optional v0 / v1 overrides)
        if v0:
            self.v1 = v0
        if v1:
            self.v0 = v1
        if camera_centers_world is None:
            raise ValueError("Missing camera_centers_world argument. Please, provide Tuple with coordinates of the camera centers "
                             "in world reference system to be added")
        self.v0 = np.concatenate((self.v0, np.asarray(self.cameras[0].C).reshape(1, -1), np.asarray(self.cameras[1].C).reshape(1, -1)))
        self.v1 = np.concatenate((self.v1, camera_centers_world))

    def triangulate_image_points(self, image_points: List[np.ndarray]) -> np.ndarray:
        triangulation = Triangulate(self.cameras, image_points)
        triangulation.triangulate_two_views()
        return triangulation.points3d

    def estimate_transformation_linear(self, estimate_scale: bool = True) -> np.ndarray:
        """absolute_orientation.py:141-154: closed-form similarity v1 ~ T v0 (moments on the device, 4x4 eigenproblem on the host)."""
        a = torch.as_tensor(np.ascontiguousarray(self.v0, dtype=np.float64)).cuda()
        b = torch.as_tensor(np.ascontiguousarray(self.v1, dtype=np.float64)).cuda()
        self.tform = similarity_from_moments(ops.helmert_moments(a, b).cpu().numpy(), estimate_scale)
        return self.tform

    def extract_params_from_T(self, T: np.ndarray = None):
        if T is None:
            T = self.tform
        t = T[:3, 3:4].squeeze()
        rot = euler_from_matrix_sxyz(T[:3, :3])
        return {"rx": rot[0], "ry": rot[1], "rz": rot[2], "tx": t[0], "ty": t[1], "tz": t[2], "m": float(1.0)}

    def estimate_transformation_least_squares(self, uncertainty: np.ndarray = None) -> np.ndarray:
        try:
            import lmfit  # noqa: F401
        except ImportError:
            raise ImportError("lmfit is not installed. Please, install it by running 'pip install lmfit'")
        raise NotImplementedError("the lmfit refinement (icepy4d/least_squares) is outside the B200 hot path (SURVEY.md §2)")

    def apply_transformation(self, T: np.ndarray = None, points3d: np.ndarray = None, camera=None) -> np.ndarray:
        """absolute_orientation.py:247-287: transforms the points (device kernel; a device tensor stays on the device) and the
        camera poses (host 4x4 products through the Camera's own pose / pose_to_extrinsics / update_extrinsics)."""
        assert not (self.v1 is None and points3d is None), \
            "Points to be transformed not found in self.v1 and not provided. Please provide a set of points to be transformed."
        if T is None:
            T = self.tform
        if points3d is None:
            points3d = self.v1
        if isinstance(points3d, torch.Tensor):
            self.v1 = ops.apply_transform(points3d.to(torch.float64), T)
        else:
            X = torch.as_tensor(np.ascontiguousarray(points3d, dtype=np.float64)).cuda()
            self.v1 = ops.apply_transform(X, T).cpu().numpy()
        for cam in (self.cameras if camera is None else [camera]):
            pose = T @ cam.pose
            cam.update_extrinsics(cam.pose_to_extrinsics(pose))
        return self.v1
