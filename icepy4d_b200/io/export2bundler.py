"""Bundler .out exporter — drop-in for icepy4d/io/export2bundler.py:90-172 (the on-disk format the epoch's
features / points / cameras are handed to Metashape in).  Same file, byte for byte, as the reference writes for the same inputs
(tests/test_host_logic.py against tests/golden/containers.npz); the per-point Python string formatting loop of the reference is
replaced by one vectorised formatting pass."""
from __future__ import annotations

import logging
import math
from pathlib import Path
from typing import Dict, Union

import numpy as np


def _bundler_extrinsics(cam):
    """The camera rotated by 180 degrees about its x axis (Bundler looks down -z): pose' = pose @ euler_matrix(pi, 0, 0), then
    back to extrinsics — the same sequence of f64 products as export2bundler.py:118-124 / core/camera.py:291-331."""
    R = np.asarray(cam.R, np.float64).reshape(3, 3)
    t = np.asarray(cam.t, np.float64).reshape(3, 1)
    Rc = R.T
    C = -np.dot(Rc, t)
    pose = np.eye(4)
    pose[:3, :3], pose[:3, 3:4] = Rc, C
    si = math.sin(math.pi)
    Rx = np.identity(4)
    Rx[1, 1], Rx[1, 2], Rx[2, 1], Rx[2, 2] = -1.0, -si, si, -1.0      # thirdparty/transformations.py:1103-1123, axes "sxyz"
    pose = pose @ Rx
    R2 = pose[:3, :3].T
    t2 = -R2 @ pose[:3, 3:4]
    return R2, t2.squeeze()


def write_bundler_out(export_dir: Union[str, Path], fname: str, images: Dict[str, object], cameras: Dict[str, object],
                      features: Dict[str, object], points) -> bool:
    """images: {cam: object with .path (or a path)}, cameras: {cam: camera with K, dist, R, t, width, height},
    features: {cam: Features}, points: Points.  Writes <export_dir>/<fname>.out and im_list.txt."""
    logging.info("Exporting results in Bundler format...")
    cams = list(cameras.keys())
    export_dir = Path(export_dir)
    export_dir.mkdir(parents=True, exist_ok=True)
    num_pts = len(features[cams[0]])
    w, h = cameras[cams[0]].width, cameras[cams[0]].height
    lines = ["# Bundle file v0.3", f"{len(cams)} {num_pts}"]
    for cam in cams:
        c = cameras[cam]
        R, t = _bundler_extrinsics(c)
        K, dist = np.asarray(c.K), np.asarray(c.dist).reshape(-1)
        lines.append(f"{K[1, 1]:.10f} {dist[0]:.10f} {dist[1]:.10f}")
        lines += [f"{r[0]:.10f} {r[1]:.10f} {r[2]:.10f}" for r in R]
        lines.append(f"{t[0]:.10f} {t[1]:.10f} {t[2]:.10f}")
    xyz = points.to_numpy()
    col = points.colors_to_numpy(as_uint8=True)
    im = []
    for cam in cams[:2]:
        m = features[cam].kpts_to_numpy()                    # f32 arithmetic up to the + [0.5, -0.5], which promotes to f64 (:147-152)
        m[:, 0] = m[:, 0] - w / 2
        m[:, 1] = h / 2 - m[:, 1]
        im.append(m + np.array([0.5, -0.5]))
    sx = [repr(v) for v in xyz.reshape(-1).astype(np.float64).tolist()]     # f"{np.float32}" formats the value as a Python float
    a, b = im
    for i in range(num_pts):
        lines.append(f"{sx[3 * i]} {sx[3 * i + 1]} {sx[3 * i + 2]}")
        lines.append(f"{col[i, 0]} {col[i, 1]} {col[i, 2]}")
        lines.append(f"2 0 {i} {a[i, 0]:.4f} {a[i, 1]:.4f} 1 {i} {b[i, 0]:.4f} {b[i, 1]:.4f}")
    with open(export_dir / f"{fname}.out", "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(export_dir / "im_list.txt", "w") as f:
        for cam in cams:
            f.write(f"{getattr(images[cam], 'path', images[cam])}\n")
    logging.info("Export to Bundler format completed successfully.")
    return True
