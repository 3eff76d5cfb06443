"""TEST INFRASTRUCTURE — not part of the product.

Import shims that let the *unmodified* reference (`/root/reference/src/icepy4d`, or the copy `oracle/stage_ref.py`
stages into the git-ignored oracle/_ref/ so that it can travel to the GPU box) run on CPU.
Used ONLY by `oracle/make_golden.py`, by `bench.py --impl reference` / the cpu_baseline leg, and by `-m "not gpu"`
tests that are skipped when no reference tree is available.  Nothing from the reference is
copied: the shims are stubs for third-party modules the container lacks (easydict, matplotlib, kornia,
exifread, open3d, laspy) and a context manager that neutralises checkpoint loading so seeded weights
(`icepy4d_b200.weights`) can be installed instead.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")      # oracle/stage_ref.py (git-ignored)
REFERENCE_SRC = os.environ.get("ICEPY4D_REFERENCE_SRC",
                               "/root/reference/src" if os.path.isdir("/root/reference/src/icepy4d") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "icepy4d"))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        d = _Dummy(f"{self.__name__}.{name}")
        setattr(self, name, d)
        return d

    def __call__(self, *a, **k):
        return None


def install_shims() -> None:
    import torch

    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    for name in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = _Dummy(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    if "kornia" not in sys.modules:
        k = _Dummy("kornia")
        kc = _Dummy("kornia.color")
        kf = _Dummy("kornia.feature")

        def rgb_to_grayscale(img):
            w = img.new_tensor([0.299, 0.587, 0.114]).view(-1, 1, 1)
            return (img * w).sum(-3, keepdim=True)

        def grayscale_to_rgb(img):
            return torch.cat([img, img, img], dim=-3)

        kc.rgb_to_grayscale = rgb_to_grayscale
        kc.grayscale_to_rgb = grayscale_to_rgb
        k.color = kc
        k.feature = kf
        sys.modules["kornia"] = k
        sys.modules["kornia.color"] = kc
        sys.modules["kornia.feature"] = kf
    for name in ("exifread", "open3d", "laspy", "h5py", "lmfit", "pydegensac_absent"):
        if name not in sys.modules:
            sys.modules[name] = _Dummy(name)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


@contextlib.contextmanager
def no_checkpoint_loading():
    """Make `torch.load`, `torch.hub.load_state_dict_from_url` and `Module.load_state_dict` no-ops
    while reference nets are constructed (their checkpoints are missing)."""
    import torch
    from torch import nn

    o_load, o_hub, o_lsd = torch.load, torch.hub.load_state_dict_from_url, nn.Module.load_state_dict
    torch.load = lambda *a, **k: {}
    torch.hub.load_state_dict_from_url = lambda *a, **k: {}
    nn.Module.load_state_dict = lambda self, *a, **k: None
    try:
        yield
    finally:
        torch.load, torch.hub.load_state_dict_from_url, nn.Module.load_state_dict = o_load, o_hub, o_lsd


def build_reference_superpoint_sg(conf: dict, state):
    """Reference SuperPoint (SuperGlue flavour) with seeded weights."""
    install_shims()
    from icepy4d.thirdparty.SuperGlue.models.superpoint import SuperPoint

    with no_checkpoint_loading():
        net = SuperPoint(conf)
    net.load_state_dict(state)
    return net.eval()


def build_reference_superglue(conf: dict, state):
    install_shims()
    from icepy4d.thirdparty.SuperGlue.models.superglue import SuperGlue

    with no_checkpoint_loading():
        net = SuperGlue(conf)
    net.load_state_dict(state)
    return net.eval()


def build_reference_superpoint_lg(state, **conf):
    install_shims()
    from icepy4d.thirdparty.LightGlue.lightglue.superpoint import SuperPoint

    with no_checkpoint_loading():
        net = SuperPoint(**conf)
    net.load_state_dict(state)
    return net.eval()


def build_reference_lightglue(state, **conf):
    install_shims()
    from icepy4d.thirdparty.LightGlue.lightglue.lightglue import LightGlue

    with no_checkpoint_loading():
        net = LightGlue(features="superpoint", **conf)
    net.load_state_dict(state, strict=False)
    return net.eval()
