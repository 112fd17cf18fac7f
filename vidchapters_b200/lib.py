"""ctypes binding of libvidchap.so (the C ABI in include/vidchap.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C vidchapters_b200/csrc`.  Loading fails loudly
if it is missing: there is no CPU or PyTorch fallback for any op (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvidchap.so")
ABI_VERSION = 1


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_fp32", C.c_int32), ("atomic", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("act", C.c_int32),
        ("pre_out", C.c_void_p),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64),
        ("alpha", C.c_float),
        ("splits", C.c_int32),
        ("tile_n", C.c_int32),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vidchapters_b200/csrc`). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.vc_last_error.restype = C.c_char_p
    lib.vc_version.restype = C.c_int
    if lib.vc_version() != ABI_VERSION:
        raise RuntimeError(f"libvidchap ABI {lib.vc_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise RuntimeError("libvidchap: " + load().vc_last_error().decode())
