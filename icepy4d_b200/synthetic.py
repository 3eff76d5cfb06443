"""Seeded synthetic inputs for benchmarks and parity tests.

* `stereo_pair`: blurred-noise image pair; image 1 is a crop of the same canvas shifted by a multiple of
  8 px, so SuperPoint (exactly equivariant to 8-px shifts) yields known ground-truth correspondences.
* `two_view_scene`: the cfg3 geometry — two calibrated cameras (intrinsics of the reference's
  assets/calib/cam{1,2}.txt), N correspondences with Brown distortion, pixel noise and a fraction of outliers.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np


def blurred_noise(h: int, w: int, seed: int, sigma: float = 1.5) -> np.ndarray:
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), sigma).astype(np.float32)
    img = (img - img.min()) / max(float(img.max() - img.min()), 1e-6) * 255.0
    return img.astype(np.uint8)


def stereo_pair(h: int, w: int, seed: int = 1000, shift=(16, 24), channels: int = 3):
    """Returns (image0, image1) uint8 [h,w,channels] (or [h,w] if channels == 1); image1(x,y) = image0(x+sx, y+sy)."""
    sx, sy = shift
    canvas = blurred_noise(h + sy, w + sx, seed)
    i0, i1 = canvas[:h, :w], canvas[sy:sy + h, sx:sx + w]
    if channels == 1:
        return np.ascontiguousarray(i0), np.ascontiguousarray(i1)
    return (np.ascontiguousarray(np.repeat(i0[:, :, None], channels, 2)),
            np.ascontiguousarray(np.repeat(i1[:, :, None], channels, 2)))


# intrinsics of the reference's assets/calib/cam1.txt / cam2.txt (6012 x 4008)
CAM1_K = np.array([[6621.74345720628298, 0, 3013.24420057086490], [0, 6621.74345720628298, 1943.47461466223308], [0, 0, 1]])
CAM1_DIST = np.array([-9.41830394356213407e-02, 8.55303528514532035e-02, 1.68948638308769863e-04, -8.74637609310216697e-04, 0.0])
CAM2_K = np.array([[9267.89262766209504, 0, 3053.49107994520591], [0, 9267.89262766209504, 1948.35654532114540], [0, 0, 1]])
CAM2_DIST = np.array([-8.07042713029020586e-02, 9.46617629940955385e-02, 3.31782983128223608e-04, -4.32106111976037410e-04, 0.0])
IMG_W, IMG_H = 6012, 4008


@dataclass
class SimpleCamera:
    """The three attributes of icepy4d.core.camera.Camera the path reads (camera.py:117-207)."""
    K: np.ndarray
    dist: np.ndarray
    R: np.ndarray
    t: np.ndarray

    @property
    def P(self) -> np.ndarray:
        return self.K @ np.hstack([self.R, self.t.reshape(3, 1)])

    # the extrinsics / pose helpers RelativeOrientation.estimate_pose calls on a Camera (camera.py:139-207, 263-330)
    @property
    def extrinsics(self) -> np.ndarray:
        return self.Rt_to_extrinsics(self.R, self.t)

    @property
    def pose(self) -> np.ndarray:
        return np.linalg.inv(self.extrinsics)

    @property
    def C(self) -> np.ndarray:
        return self.pose[0:3, 3:4]

    @staticmethod
    def Rt_to_extrinsics(R, t) -> np.ndarray:
        E = np.eye(4)
        E[:3, :3] = np.asarray(R, dtype=np.float64)
        E[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
        return E

    @staticmethod
    def pose_to_extrinsics(pose) -> np.ndarray:
        return np.linalg.inv(np.asarray(pose, dtype=np.float64))

    def update_extrinsics(self, extrinsics) -> None:
        self.R = np.asarray(extrinsics)[:3, :3].copy()
        self.t = np.asarray(extrinsics)[:3, 3].copy()


def _look_at(C, target):
    z = target - C
    z = z / np.linalg.norm(z)
    x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return R, -R @ C


def _project(X, cam: SimpleCamera):
    Xc = X @ cam.R.T + cam.t
    x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
    k1, k2, p1, p2, k3 = cam.dist
    r2 = x * x + y * y
    rad = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([xd * cam.K[0, 0] + cam.K[0, 2], yd * cam.K[1, 1] + cam.K[1, 2]], 1), Xc[:, 2]


def two_view_scene(n: int = 200_000, seed: int = 7, noise_px: float = 0.3, outlier_frac: float = 0.4, distortion: bool = True):
    """Returns dict(cams=[cam0, cam1], pts0 [n,2] f32, pts1 [n,2] f32, X [n,3] f64, inlier [n] bool).
    distortion=False gives ideal pinhole cameras (relative-orientation tests work on undistorted coordinates)."""
    rng = np.random.default_rng(seed)
    d0, d1 = (CAM1_DIST, CAM2_DIST) if distortion else (np.zeros(5), np.zeros(5))
    cam0 = SimpleCamera(CAM1_K, d0, np.eye(3), np.zeros(3))
    R1, t1 = _look_at(np.array([250.0, 0.0, 30.0]), np.array([100.0, 0.0, 600.0]))
    cam1 = SimpleCamera(CAM2_K, d1, R1, t1)
    Xs, p0s, p1s = [], [], []
    have = 0
    while have < n:
        X = np.stack([rng.uniform(-120, 320, 2 * n), rng.uniform(-120, 120, 2 * n), rng.uniform(450, 800, 2 * n)], 1)
        a, za = _project(X, cam0)
        b, zb = _project(X, cam1)
        ok = (za > 0) & (zb > 0) & (a[:, 0] > 0) & (a[:, 0] < IMG_W) & (a[:, 1] > 0) & (a[:, 1] < IMG_H) \
            & (b[:, 0] > 0) & (b[:, 0] < IMG_W) & (b[:, 1] > 0) & (b[:, 1] < IMG_H)
        Xs.append(X[ok]); p0s.append(a[ok]); p1s.append(b[ok])
        have += int(ok.sum())
    X = np.concatenate(Xs)[:n]
    p0 = np.concatenate(p0s)[:n] + rng.normal(0, noise_px, (n, 2))
    p1 = np.concatenate(p1s)[:n] + rng.normal(0, noise_px, (n, 2))
    inl = np.ones(n, dtype=bool)
    n_out = int(round(outlier_frac * n))
    if n_out:
        idx = rng.choice(n, n_out, replace=False)
        p1[idx] = np.stack([rng.uniform(0, IMG_W, n_out), rng.uniform(0, IMG_H, n_out)], 1)
        inl[idx] = False
    return {"cams": [cam0, cam1], "pts0": p0.astype(np.float32), "pts1": p1.astype(np.float32), "X": X, "inlier": inl}
