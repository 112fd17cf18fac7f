"""ctypes binding of libvidchap.so (the C ABI in include/vidchap.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C vidchapters_b200/csrc`.  Loading fails loudly
if it is missing: there is no CPU or PyTorch fallback for any op (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvidchap.so")
ABI_VERSION = 9


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
        ("alpha_dev", C.c_void_p),
        ("splits", C.c_int32),
        ("tile_n", C.c_int32),
        ("drop_seed", C.c_uint32), ("drop_p16", C.c_uint32),
        ("ce_labels", C.c_void_p), ("ce_stats", C.c_void_p), ("ce_zy", C.c_void_p), ("ce_lse", C.c_void_p),
        ("ce_nvalid", C.c_void_p), ("ce_smoothing", C.c_float),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
        ("q_col", C.c_int32), ("k_col", C.c_int32), ("v_col", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("head_dim", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("lse2", C.c_void_p),
        ("bias_rel", C.c_void_p),
        ("kmask", C.c_void_p),
        ("causal", C.c_int32),
        ("scale", C.c_float),
        ("drop_seed", C.c_uint32), ("drop_p16", C.c_uint32),
        ("q_offset", C.c_int32), ("q_offset_dev", C.c_void_p), ("kv_batch_rows", C.c_int32),
        ("bias_zero", C.c_int32), ("bias_len", C.c_int32),
        ("q_like_k", C.c_int32),
        ("kv_batch_div", C.c_int32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", AttnArgs),
        ("dout", C.c_void_p), ("ld_do", C.c_int64), ("do_col", C.c_int32),
        ("delta", C.c_void_p),
        ("dq_acc", C.c_void_p), ("ld_dq", C.c_int64),
        ("dk", C.c_void_p), ("ld_dk", C.c_int64), ("dk_col", C.c_int32),
        ("dv", C.c_void_p), ("ld_dv", C.c_int64), ("dv_col", C.c_int32),
        ("dbias_rel", C.c_void_p),
        ("bucket_lut", C.c_void_p),
    ]


P, I, I64, F, U = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32
SIGNATURES = {
    "vc_gemm_bf16": [C.POINTER(GemmArgs), P],
    "vc_attn_fwd": [C.POINTER(AttnArgs), P],
    "vc_attn_bwd": [C.POINTER(AttnBwdArgs), P],
    "vc_norm_fwd": [I, P, P, P, P, P, P, P, I, I, F, F, I, I, I, U, U, P],
    "vc_norm_bwd": [I, P, I, P, P, P, P, P, P, I, P, P, I, I, F, I, I, I, U, U, U, U, P],
    "vc_embed_fwd": [P, P, P, I, I, I, U, U, P],
    "vc_embed_bwd": [P, P, P, I, I, I, U, U, P],
    "vc_prepare_targets": [P, P, P, P, I, I, I64, P],
    "vc_bias_expand": [P, P, P, I, I, P],
    "vc_bias_fold": [P, P, P, I, I, P],
    "vc_add_pos": [P, P, P, I, I, I, I, U, U, P],
    "vc_add_pos_bwd": [P, P, I, I, I, I, U, U, P],
    "vc_cross_entropy": [P, I64, P, P, F, P, P, I64, I, I, P],
    "vc_ce_combine": [P, I, P, P, P, F, I, P, P, I, P],
    "vc_colsum_bf16": [P, I64, P, I, I, P],
    "vc_cast_f32_bf16": [P, I64, P, I64, I, I, F, P],
    "vc_copy_rows_bf16": [P, P, I, I, I, I, I, P],
    "vc_sumsq": [P, I64, P, P],
    "vc_adam_step": [P, P, P, P, P, I64, F, F, F, F, I, P, F, F, P],
    "vc_renorm_time_tokens": [P, P, I, I, I, P, P],
    "vc_cast_flat_bf16": [P, P, I64, P],
    "vc_kv_append": [P, I64, P, I, I, I, P, P],
    "vc_greedy_next": [P, I64, I, P, P, P, I, P, I64, I64, I, P],
    "vc_step_advance": [P, P],
    "vc_decode_linear": [P, I64, I, P, F, F, P, I64, P, I64, I, P, I64, I, I, I, I, P],
    "vc_beam_topk": [P, I64, I, P, I, I, P, P, P, P],
    "vc_kv_reorder": [P, P, P, I, I, I, I, P],
    "vc_set_dropout_salt": [P],
    "vc_debug_set_trace": [P],
    "vc_version": [],
    "vc_last_error": [],
    "vc_device_check": [],
}

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
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what include/vidchap.h declares
        fn.argtypes = argtypes
        if name != "vc_last_error":
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise RuntimeError("libvidchap: " + load().vc_last_error().decode())
