"""ctypes binding of the C ABI in include/icepy4d_b200.h.

Prototypes are parsed from the header itself, so Python can never drift from the ABI.  There is NO fallback: if the
shared library is missing or a call fails, an exception is raised (the product path must fail loudly without its
CUDA extension).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "icepy4d_b200.h")
LIB_PATH = os.path.join(_HERE, "_lib", "libicepy4d_b200.so")

_PROTO_RE = re.compile(r"^(int|size_t|const char\*)\s+(i4d_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)


class NativeError(RuntimeError):
    pass


def _ctype(decl: str):
    decl = decl.strip()
    if decl == "void":
        return None
    if "*" in decl:
        return ctypes.c_void_p
    base = re.sub(r"\b(const|unsigned)\b", "", decl).split()
    unsigned = "unsigned" in decl
    t = " ".join(base[:-1]) if len(base) > 1 else base[0]
    table = {"int": ctypes.c_uint if unsigned else ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
             "size_t": ctypes.c_size_t, "long long": ctypes.c_ulonglong if unsigned else ctypes.c_longlong}
    if t not in table:
        raise NativeError(f"unsupported C type in header: {decl!r}")
    return table[t]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """name -> (return type, [argument declarations])"""
    with open(path) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    out = {}
    for ret, name, args in _PROTO_RE.findall(src):
        out[name] = (ret, [a.strip() for a in args.replace("\n", " ").split(",")])
    return out


_lib = None
_protos = None


def lib() -> ctypes.CDLL:
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} not found — build it with `python -m icepy4d_b200.build` "
                          "(there is no CPU or PyTorch fallback for this path)")
    l = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, args) in _protos.items():
        fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p}[ret]
        fn.argtypes = [t for t in (_ctype(a) for a in args) if t is not None]
    _lib = l
    return l


def last_error() -> str:
    return lib().i4d_last_error().decode()


def _ptr(x):
    """torch tensor / numpy array / int / None -> raw address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


def call(name: str, *args):
    """Call an int-returning entry point; pointers may be tensors/arrays; raises NativeError on failure."""
    l = lib()
    fn = getattr(l, name)
    if len(args) != len(fn.argtypes):
        raise TypeError(f"{name}: expected {len(fn.argtypes)} arguments, got {len(args)}")
    conv = [(_ptr(a) if t is ctypes.c_void_p else a) for a, t in zip(args, fn.argtypes)]
    rc = fn(*conv)
    if fn.restype is ctypes.c_int and rc != 0:
        raise NativeError(f"{name} failed ({rc}): {last_error()}")
    return rc


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


# ---- kernel-launch accounting (the claim reported by bench.py as `gpu_launches`) ----------------------------
LAUNCHES = 0
_PER_CALL = {"i4d_sp_score_map": 1, "i4d_sp_conv1a_relu": 1, "i4d_sp_conv1ab_tc": 1, "i4d_maxpool2x2_nhwc": 1, "i4d_sp_nms_candidates": 1, "i4d_sp_select_topk": 10, "i4d_sp_sample_descriptors": 1,
             "i4d_gemm_f32": 1, "i4d_attention_f32": 1, "i4d_layernorm_gelu": 1, "i4d_lg_posenc": 1, "i4d_lg_rotary": 1,
             "i4d_sg_kenc_input": 1, "i4d_set_sinkhorn_mode": 0, "i4d_row_lse": 1, "i4d_col_lse": 2, "i4d_lg_assign": 14, "i4d_undistort_points": 1,
             "i4d_triangulate_iterative_ls": 1, "i4d_triangulate_dlt": 1, "i4d_tile_to_gray_f32": 1, "i4d_pyr_down_u8": 1, "i4d_pyr_up_u8": 1, "i4d_essential_pose": 4, "i4d_interpolate_point_colors": 1}


def _count(name: str, args) -> int:
    if name in _PER_CALL:
        return _PER_CALL[name]
    if name in ("i4d_sinkhorn", "i4d_sg_assign"):
        n, ld, iters = int(args[2]), int(args[3]), int(args[5])
        fused = ld % 4 == 0 and ld >= (n + 3) // 4 * 4 and 64 <= n <= 8192   # one persistent cooperative launch for all iterations
        return (1 + (1 if n % 4 else 0) if fused else 5 * iters) + (5 if name == "i4d_sg_assign" else 0)
    if name == "i4d_essential_ransac":
        return 2 + 3 * min(64, -(-int(args[5]) // 512))
    if name == "i4d_fundamental_ransac":
        rounds = min(24, -(-int(args[5]) // 4096))
        return 1 + 5 * rounds + 1 + 2 * int(args[8]) + 2
    return 1


_raw_call = call


def call(name: str, *args):  # noqa: F811
    global LAUNCHES
    rc = _raw_call(name, *args)
    LAUNCHES += _count(name, args)
    return rc
