"""CPU check of the five-point essential-matrix solver (icepy4d_b200/csrc/five_point.cuh is plain C++: the same code the CUDA
RANSAC of csrc/essential.cu runs one sample per thread is built here with g++ and driven through ctypes).  It replaces the
minimal solver behind cv2.findEssentialMat at /root/reference/src/icepy4d/sfm/geometry.py:63-65; cv2 itself is the cross-check."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = """
#include "five_point.cuh"
extern "C" int fp_solve(const double* x0, const double* x1, double* E) {
  return fivept::solve(reinterpret_cast<const double (*)[2]>(x0), reinterpret_cast<const double (*)[2]>(x1),
                       reinterpret_cast<double (*)[9]>(E));
}
"""


@pytest.fixture(scope="module")
def fp(tmp_path_factory):
    d = tmp_path_factory.mktemp("fp")
    (d / "fp_host.cpp").write_text(SRC)
    so = str(d / "libfp.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "icepy4d_b200", "csrc"), "-o", so, str(d / "fp_host.cpp")],
                   check=True)
    lib = ctypes.CDLL(so)
    lib.fp_solve.argtypes = [ctypes.c_void_p] * 3
    lib.fp_solve.restype = ctypes.c_int
    return lib


def _rot(rng, ang=0.3):
    w = rng.normal(size=3)
    w /= np.linalg.norm(w)
    a = rng.uniform(-ang, ang)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def _scene(rng, planar):
    R, t = _rot(rng), rng.normal(size=3)
    t /= np.linalg.norm(t)
    X = rng.uniform(-1, 1, size=(5, 3)) + np.array([0, 0, 4.0])
    if planar:
        X[:, 2] = 4.0 + 0.3 * X[:, 0]
    x0 = X[:, :2] / X[:, 2:]
    Xc = X @ R.T + t
    x1 = Xc[:, :2] / Xc[:, 2:]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Et = tx @ R
    return np.ascontiguousarray(x0), np.ascontiguousarray(x1), Et / np.linalg.norm(Et)


def _solve(fp, x0, x1):
    E = np.zeros((10, 9))
    n = fp.fp_solve(x0.ctypes.data, x1.ctypes.data, E.ctypes.data)
    return E[:n].reshape(n, 3, 3)


def test_true_essential_matrix_is_among_the_solutions(fp):
    rng = np.random.default_rng(0)
    found, total, violations, nsol = 0, 600, 0, 0
    for trial in range(total):
        x0, x1, Et = _scene(rng, planar=trial % 5 == 0)
        sols = _solve(fp, x0, x1)
        nsol += len(sols)
        h0, h1 = np.c_[x0, np.ones(5)], np.c_[x1, np.ones(5)]
        for Ek in sols:
            res = np.abs(np.einsum("ki,ij,kj->k", h1, Ek, h0)).max()                       # the five epipolar constraints
            cub = np.abs(2 * Ek @ Ek.T @ Ek - np.trace(Ek @ Ek.T) * Ek).max()              # essential-manifold constraint
            violations += int(res > 1e-7 or cub > 1e-5)
            assert abs(np.linalg.norm(Ek) - 1) < 1e-9
        found += int(any(min(np.abs(Ek - Et).max(), np.abs(Ek + Et).max()) < 1e-6 for Ek in sols))
    assert found >= 0.97 * total, found                # ill-conditioned samples may lose a root; a RANSAC just draws again
    assert violations <= 0.02 * nsol, (violations, nsol)
    assert 2.0 < nsol / total <= 10.0


def test_same_solutions_as_opencv(fp):
    """cv2.findEssentialMat on exactly five points returns all solutions of the minimal problem stacked as [3k, 3]."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    agree, total = 0, 100
    for _ in range(total):
        x0, x1, _ = _scene(rng, planar=False)
        ours = _solve(fp, x0, x1)
        Ecv, _ = cv2.findEssentialMat(x0, x1, np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=1e-3)
        if Ecv is None:
            continue
        theirs = [e / np.linalg.norm(e) for e in np.split(Ecv, len(Ecv) // 3)]
        ok = all(any(min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-5 for a in ours) for b in theirs)
        agree += int(ok)
    assert agree >= 0.9 * total, agree
