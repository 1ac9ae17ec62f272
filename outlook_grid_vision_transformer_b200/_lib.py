"""ctypes binding of libogvit.so (include/ogv.h).  There is no fallback: if the shared library is
missing and cannot be built in-tree, importing this module's `lib()` raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libogvit.so"

F32, BF16 = 0, 1
ACT = {None: 0, "none": 0, "gelu": 1, "silu": 2, "sigmoid": 3, "relu": 4, "mul": 5}
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2


class GemmArgs(Structure):
    """Mirror of `struct ogv_gemm_args` (include/ogv.h)."""

    _fields_ = [
        ("A", c_void_p), ("a_rs", c_longlong), ("a_cs", c_longlong),
        ("B", c_void_p), ("b_rs", c_longlong), ("b_cs", c_longlong),
        ("D", c_void_p), ("ldd", c_longlong),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("in_dtype", c_int), ("out_dtype", c_int),
        ("bias", c_void_p),
        ("pre_out", c_void_p), ("ld_pre", c_longlong),
        ("act", c_int),
        ("dact_src", c_void_p), ("ld_dact", c_longlong), ("dact", c_int),
        ("row_scale", c_void_p), ("rows_per_scale", c_int),
        ("residual", c_void_p), ("ld_res", c_longlong),
        ("accumulate", c_int), ("split_k", c_int),
        ("col_sum", c_void_p), ("col_sumsq", c_void_p),
        ("pre_out_grad", c_int),
        ("row_sum", c_void_p),
    ]


_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float

# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/ogv.h exactly.
SIGNATURES = {
    "ogv_version": [],
    "ogv_last_error": [],
    "ogv_sm_count": [],
    "ogv_gemm": [POINTER(GemmArgs), _I, _P],
    "ogv_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I, _P],
    "ogv_nhwc_to_nchw": [_P, _P, _I, _I, _I, _I, _P],
    "ogv_cast_transpose": [_P, _P, _L, _P, _L, _I, _I, _I, _P],
    "ogv_cast_batch": [_P, _I, _I, _P],
    "ogv_split3": [_P, _L, _P, _L, _L, _L, _I, _I, _P],
    "ogv_rowscale": [_P, _P, _P, _L, _I, _I, _I, _P],
    "ogv_rowscale_colsum": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "ogv_colsum": [_P, _L, _P, _L, _I, _I, _P],
    "ogv_mul_dact": [_P, _P, _P, _L, _I, _I, _P],
    "ogv_add": [_P, _P, _P, _L, _I, _P],
    "ogv_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _L, _I, _F, _I, _P],
    "ogv_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ogv_outlook_core_fwd": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_outlook_core_bwd": [_P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_colstats": [_P, _L, _P, _P, _L, _I, _I, _P],
    "ogv_bn_finalize": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _I, _P],
    "ogv_bn_apply": [_P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ogv_bn_bwd_reduce": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ogv_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ogv_se_mlp_supported": [_I, _I, _I],
    "ogv_se_mlp_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ogv_se_mlp_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ogv_im2col3x3": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_im2col3x3_vec": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_col2im3x3_vec": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_conv3x3_supported": [_I, _I, _I, _I, _I, _I],
    "ogv_conv3x3_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_conv3x3_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_bn_act_apply": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "ogv_bn_act_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "ogv_bn_act_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "ogv_dwconv_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_se_pool": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ogv_bn_act_gate": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ogv_se_bwd_reduce": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ogv_mbconv_bwd_stats": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ogv_mbconv_bn2_finalize": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ogv_dw_bn2_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ogv_dwconv_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ogv_grid_attn_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ogv_grid_attn_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ogv_grid_attn_probs": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ogv_adamw": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _F, _F, _P],
    "ogv_xent_fwd": [_P, _L, _P, _I, _I, _F, _I, _P, _P, _P],
    "ogv_xent_bwd": [_P, _L, _P, _P, _P, _I, _I, _F, _I, _P, _L, _P],
    "ogv_sumsq": [_P, _L, _P, _P],
    "ogv_adamw_flat": [_P, _P, _P, _P, _P, _L, _P, _P, _P, _F, _F, _F, _F, _P, _P],
    "ogv_train_metrics": [_P, _L, _P, _I, _I, _P, _P, _P],
    "ogv_mlp_fused_supported": [_I, _I],
    "ogv_mlp_fwd": [_P, _L, _P, _P, _P, _P, _P, _L, _P, _I, _P, _L, _L, _I, _I, _I, _P],
    "ogv_mlp_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _P, _I, _P, _P, _P, _L, _L, _I, _I, _I, _P],
}
_RESTYPES = {"ogv_last_error": c_char_p}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (building in-tree on first use if a toolchain is present) and return libogvit.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() or os.environ.get("OGV_REBUILD") == "1":
        from . import build as _build  # nvcc cross-compiles without a GPU

        _build.build(force=os.environ.get("OGV_REBUILD") == "1")  # file-locked: one builder across ranks
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension is required, there is no CPU fallback")
    # OGV_LIB: load another build of the SAME library (A/B measurements of a kernel change on one box)
    handle = ctypes.CDLL(os.environ.get("OGV_LIB") or str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = handle
    return handle


def last_error() -> str:
    msg = lib().ogv_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        err = last_error()
        exc = ValueError if rc == -1 else (NotImplementedError if rc == -3 else RuntimeError)
        raise exc(f"libogvit {what} failed (code {rc}): {err}")
