"""ctypes binding of libmt_b200.so (the C ABI declared in include/mt_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Build it with ``python -m musicgeneration_b200.build`` (or
``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MT_LIB_PATH") or os.path.join(_HERE, "libmt_b200.so")     # (override: A/B timing of two builds)

MT_F32, MT_BF16, MT_F16 = 0, 1, 2
MT_F16_BF16 = 3        # mt_rga_fwd / mt_rga_bwd_ws only: f16 q/k/v/E, bf16 O/dO/dq/dk/dv (include/mt_b200.h)
EPI_BIAS, EPI_RELU, EPI_ADD, EPI_RELU_MASK = 1, 2, 4, 8
PATH_AUTO, PATH_SIMT, PATH_TC = 0, 1, 2

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_int = C.c_int
_f = C.c_float
_u64 = C.c_uint64
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/mt_b200.h one to one
SIGNATURES = {
    "mt_version": (_int, []),
    "mt_last_error": (C.c_char_p, []),
    "mt_device_ok": (_int, []),
    "mt_embed_pos_fwd": (_int, [_p, _p, _p, _p, _p, _int, _i64, _i64, _i64, _i64, _i64, _f, _f, _u64, _u64, _p]),
    "mt_embed_pos_bwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _f, _f, _u64, _u64, _p]),
    "mt_add_ln_fwd": (_int, [_p, _int, _p, _p, _p, _p, _p, _int, _p, _p, _i64, _i64, _f, _f, _u64, _u64, _p]),
    "mt_add_ln_bwd_parts": (_i64, [_i64]),
    "mt_add_ln_bwd": (_int, [_p, _p, _int, _p, _p, _p, _p, _p, _p, _int, _p, _i64, _i64, _f, _u64, _u64, _p]),
    "mt_ln_param_grad": (_int, [_p, _p, _p, _p, _i64, _i64, _p]),
    "mt_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64, _int, _int]),
    "mt_gemm": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _int, _int, _int, _p, _sz, _p]),
    "mt_wgrad_bias": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _sz, _p]),
    "mt_colsum_workspace_bytes": (_sz, [_i64, _i64]),
    "mt_colsum": (_int, [_p, _int, _p, _i64, _i64, _i64, _p, _sz, _p]),
    "mt_cast": (_int, [_p, _int, _p, _int, _i64, _p]),
    "mt_cast2d": (_int, [_p, _int, _i64, _p, _int, _i64, _i64, _i64, _p]),
    "mt_transpose_cast": (_int, [_p, _int, _p, _int, _i64, _i64, _p]),
    "mt_rga_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p]),
    "mt_rga_weights": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _p]),
    "mt_rga_bwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p]),
    "mt_rga_bwd_ws": (_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p, _sz, _p]),
    "mt_rga_bwd_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64, _int]),
    "mt_rga_stash_bytes": (_sz, [_i64, _i64, _i64, _i64, _int]),
    "mt_rga_fwd_stash": (_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _p, _sz, _p]),
    "mt_rga_bwd_stash": (_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _p, _sz, _p, _sz, _p]),
    "mt_smooth_ce_fwd": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _f, _i32, _p]),
    "mt_smooth_ce_bwd": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _f, _i32, _p]),
    "mt_adam_step": (_int, [_p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _f, _p]),
    "mt_rga_decode": (_int, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "mt_kv_append": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "mt_sample": (_int, [_p, _p, _p, _i64, _i64, _f, _i32, _int, _p]),
    "mt_decode_embed": (_int, [_p, _i64, _p, _p, _p, _p, _p, _int, _i64, _i64, _i64, _f, _i32, _p, _i64, _p]),
    "mt_decode_kv_append": (_int, [_p, _p, _p, _p, _i64, _i32, _p, _p, _i64, _i64, _i64, _i64, _int, _p]),
    "mt_decode_attend_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64, _i64]),
    "mt_decode_attend": (_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _int, _int, _p, C.c_size_t, _p]),
    "mt_decode_sample": (_int, [_p, _p, _p, _i64, _p, _i32, _i64, _i64, _f, _i32, _int, _p]),
    "mt_decode_advance": (_int, [_p, _p]),
    "mt_decode_run_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mt_decode_run_supported": (_int, [_i64, _i64, _i64, _i64, _i64]),
    "mt_decode_run": (_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i64,
                             _i32, _p, _p, _f, _i32, _int, _p, _p, _sz, _p]),
    "mt_decode_chain": (_int, [_int]),
    "mt_window_gather": (_int, [_p, _int, _p, _p, _p, _i64, _i64, _i64, _i64, _p]),
    "mt_window_sample": (_int, [_p, _p, _i64, _i64, _u64, _u64, _p, _p, _i64, _p]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m musicgeneration_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().mt_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"libmt_b200 {what} failed (code {rc}): {last_error()}")
