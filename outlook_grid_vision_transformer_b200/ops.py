"""Tensor-level wrappers over the C-ABI (one python function per `ogv_*` entry point).

PyTorch is used here only for device memory, the current stream and dtype bookkeeping; every
computation is a libogvit kernel launched on `torch.cuda.current_stream()`.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch

from . import _lib
from ._lib import ACT, BF16, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC, F32, GemmArgs, check

Tensor = torch.Tensor


def _require_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "outlook_grid_vision_transformer_b200 runs on CUDA (sm_100a) only; got a tensor on "
                f"'{t.device}'. There is no CPU fallback.")


def dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {t.dtype}; use float32 or bfloat16")


class _Profiler:
    """Optional per-kernel timing (bench.py's roofline pass): CUDA events recorded on the launching
    stream around every library call, aggregated by kernel label; `bytes` = sum of the tensor
    arguments' sizes (each touched once = the algorithmic traffic of that launch)."""

    def __init__(self):
        self.enabled = False
        self.records = []     # (label, start_event, end_event, bytes, flops, cuda kernel)
        self.cur_bytes = 0
        self.cur_flops = 0
        self.cur_label = None
        self.cur_kernel = None  # name of the CUDA kernel behind the call when one label covers several
        self.tag = None         # (fused branch, direction, rows M, channels C) the calls belong to; set by functional.py

    def start(self):
        self.enabled, self.records, self.cur_bytes, self.cur_flops, self.cur_label = True, [], 0, 0, None
        self.cur_kernel = None
        self.tag = None

    def stop(self):
        """-> {label: dict(ms, calls, bytes, flops)} (synchronises)."""
        self.enabled = False
        torch.cuda.synchronize()
        out = {}
        self.branches = {}  # tag -> dict(ms, launches, passes): device time per fused branch (SURVEY 8(d) accounting)
        for label, e0, e1, nbytes, flops, kern, tag in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(label, dict(ms=0.0, calls=0, bytes=0, flops=0, cuda_kernel=kern or label))
            d["ms"] += ms
            d["calls"] += 1
            d["bytes"] += nbytes
            d["flops"] += flops
            if tag is not None:
                b = self.branches.setdefault(tag, dict(ms=0.0, launches=0, touched_bytes=0))
                b["ms"] += ms
                b["launches"] += 1
                b["touched_bytes"] += nbytes
        self.records = []
        return out


PROFILER = _Profiler()
LAUNCHES = 0  # library calls issued (each is at least one kernel launch)


def _p(t: Optional[Tensor]):
    if t is None:
        return None
    if PROFILER.enabled:
        PROFILER.cur_bytes += t.numel() * t.element_size()
    return ctypes.c_void_p(t.data_ptr())


def _call(name: str, *args) -> None:
    global LAUNCHES
    fn = getattr(_lib.lib(), name)
    if PROFILER.enabled:
        label = PROFILER.cur_label or name
        nbytes, flops, kern = PROFILER.cur_bytes, PROFILER.cur_flops, PROFILER.cur_kernel
        PROFILER.cur_bytes, PROFILER.cur_flops, PROFILER.cur_label, PROFILER.cur_kernel = 0, 0, None, None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILER.records.append((label, e0, e1, nbytes, flops, kern, PROFILER.tag))
    else:
        rc = fn(*args)
    LAUNCHES += 1
    check(rc, name)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: Optional[Tensor], name: str) -> None:
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous float32 tensor")


def _rows(t: Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a 2-D tensor with unit column stride, got {tuple(t.shape)} / {t.stride()}")


# ------------------------------------------------------------------------------------------- GEMM
# fp32 compute mode: run the product on the tcgen05 engine as three bf16 planes per operand (error ~2^-17 per
# product, three times the tensor work of a bf16 GEMM, still memory bound) instead of the FFMA engine.
# OGV_FP32_EXACT=1 keeps every fp32 product on the FFMA engine.
FP32_ON_TENSOR_CORES = os.environ.get("OGV_FP32_EXACT", "0") != "1"


def _split3(t: Tensor, pattern: int) -> Optional[Tensor]:
    """2-D fp32 operand [R, K] of a GEMM (K = reduction axis) -> bf16 operand with a 3x longer reduction axis and
    the same majorness, or None when the layout cannot be expressed."""
    R, K = t.shape
    if t.data_ptr() % 16:
        return None
    if t.stride(1) == 1 and K % 8 == 0 and t.stride(0) % 4 == 0:      # K-major: planes side by side -> [R, 3K]
        dst = torch.empty((R, 3 * K), device=t.device, dtype=torch.bfloat16)
        _call("ogv_split3", _p(t), t.stride(0), _p(dst), 3 * K, K, R, K, pattern, _stream())
        return dst
    if t.stride(0) == 1 and R % 8 == 0 and t.stride(1) % 4 == 0:      # MN-major view of a row-major [K, R] tensor: planes stacked -> [3K, R]
        base = torch.empty((3 * K, R), device=t.device, dtype=torch.bfloat16)
        _call("ogv_split3", _p(t), t.stride(1), _p(base), R, K * R, K, R, pattern, _stream())
        return base.t()
    return None


def gemm(A: Tensor, B: Tensor, D: Tensor, *, bias: Optional[Tensor] = None, pre_out: Optional[Tensor] = None,
         act: Optional[str] = None, dact_src: Optional[Tensor] = None, dact: Optional[str] = None,
         row_scale: Optional[Tensor] = None, rows_per_scale: int = 1, residual: Optional[Tensor] = None,
         accumulate: bool = False, split_k: int = 1, engine: int = ENGINE_AUTO, col_sum: Optional[Tensor] = None,
         col_sumsq: Optional[Tensor] = None, pre_out_grad: bool = False, row_sum: Optional[Tensor] = None) -> Tensor:
    """D[m,n] = epi(sum_k A[m,k] * B[n,k]); A:[M,K], B:[N,K] (any strides), D:[M,N] row-major.
    col_sum / col_sumsq (fp32 [N]) are accumulated (+=) with the per-column sum / sum of squares of D.
    row_sum (fp32 [M], accumulate mode): += sum_k A[m,k] -- the bias gradient when A is the transposed output gradient."""
    _require_cuda(A, B, D, bias, pre_out, dact_src, row_scale, residual)
    if A.dim() != 2 or B.dim() != 2 or D.dim() != 2:
        raise ValueError("gemm operands must be 2-D")
    M, K = A.shape
    N, K2 = B.shape
    if K != K2 or tuple(D.shape) != (M, N):
        raise ValueError(f"gemm shape mismatch A{tuple(A.shape)} B{tuple(B.shape)} D{tuple(D.shape)}")
    if A.dtype != B.dtype:
        raise TypeError("gemm: A and B must share a dtype")
    _rows(D, "D")
    for name, t in (("pre_out", pre_out), ("dact_src", dact_src), ("residual", residual)):
        if t is not None:
            _rows(t, name)
            if t.dtype != D.dtype or tuple(t.shape) != (M, N):
                raise ValueError(f"gemm: {name} must match D in dtype and shape")
    _f32(bias, "bias")
    _f32(row_scale, "row_scale")
    _f32(col_sum, "col_sum")
    _f32(col_sumsq, "col_sumsq")
    _f32(row_sum, "row_sum")
    if row_sum is not None and (not accumulate or row_sum.numel() != M):
        raise ValueError("gemm: row_sum needs accumulate=True and M elements")
    if (FP32_ON_TENSOR_CORES and engine == ENGINE_AUTO and A.dtype == torch.float32 and D.dtype == torch.float32
            and K >= 8 and M >= 64 and not (accumulate and col_sum is not None)):
        A3, B3 = _split3(A, 0), _split3(B, 1)
        if A3 is not None and B3 is not None:
            if row_sum is not None:  # the three-plane operand would sum its planes: reduce the fp32 operand itself
                colsum(A.t(), row_sum)
                row_sum = None
            gemm(A3, B3, D, bias=bias, pre_out=pre_out, act=act, dact_src=dact_src, dact=dact,
                 row_scale=row_scale, rows_per_scale=rows_per_scale, residual=residual, accumulate=accumulate,
                 split_k=split_k, engine=ENGINE_AUTO, pre_out_grad=pre_out_grad)
            # column statistics of an fp32 output: one more streaming pass (the fused form needs a bf16 output)
            if col_sum is not None and col_sumsq is not None:
                colstats(D, col_sum, col_sumsq)
            elif col_sum is not None:
                colsum(D, col_sum)
            return D
    a = GemmArgs()
    a.A, a.a_rs, a.a_cs = A.data_ptr(), A.stride(0), A.stride(1)
    a.B, a.b_rs, a.b_cs = B.data_ptr(), B.stride(0), B.stride(1)
    a.D, a.ldd = D.data_ptr(), D.stride(0)
    a.M, a.N, a.K = M, N, K
    a.in_dtype, a.out_dtype = dtype_code(A), dtype_code(D)
    a.bias = bias.data_ptr() if bias is not None else None
    a.pre_out, a.ld_pre = (pre_out.data_ptr(), pre_out.stride(0)) if pre_out is not None else (None, 0)
    a.act = ACT[act]
    a.dact_src, a.ld_dact = (dact_src.data_ptr(), dact_src.stride(0)) if dact_src is not None else (None, 0)
    a.dact = ACT[dact]
    a.row_scale = row_scale.data_ptr() if row_scale is not None else None
    a.rows_per_scale = int(rows_per_scale)
    a.residual, a.ld_res = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
    a.accumulate, a.split_k = int(accumulate), int(split_k)
    a.col_sum = col_sum.data_ptr() if col_sum is not None else None
    a.col_sumsq = col_sumsq.data_ptr() if col_sumsq is not None else None
    a.pre_out_grad = int(bool(pre_out_grad))
    a.row_sum = row_sum.data_ptr() if row_sum is not None else None
    if PROFILER.enabled:
        extra = sum(1 for t in (pre_out, dact_src, residual) if t is not None)
        PROFILER.cur_bytes = A.element_size() * (M * K + N * K) + D.element_size() * M * N * (1 + extra)
        PROFILER.cur_flops = 2 * M * N * K
        kind = "wgrad" if accumulate else ("bwd" if dact_src is not None else "fwd")
        flags = "".join(ch for ch, t in (("b", bias), ("p", pre_out), ("a", act), ("s", row_scale), ("r", residual)) if t is not None)
        PROFILER.cur_label = f"ogv_gemm[{kind} {M}x{N}x{K} {'bf16' if A.dtype == torch.bfloat16 else 'f32'} {flags}]"
        if A.dtype == torch.bfloat16 and engine != ENGINE_SIMT:
            bn = 64 if N <= 64 else (128 if N <= 128 else 256)
            PROFILER.cur_kernel = f"gemm_tc_kernel<{bn}, {'__nv_bfloat16' if D.dtype == torch.bfloat16 else 'float'}>"
        else:
            PROFILER.cur_kernel = "gemm_simt_kernel"
    _call("ogv_gemm", ctypes.byref(a), engine, _stream())
    return D


class SideStream:
    """Where the train-step executor sends work that only produces PARAMETER gradients (weight-gradient GEMMs): nothing in
    the rest of backward depends on it, so it runs on a second stream next to the dgrad chain and fills the SMs the
    latency-bound small-stage kernels leave idle.  `keep` holds the operands until the executor joins the stream (the
    caching allocator must not hand their blocks to later main-stream allocations while the side stream still reads them)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.keep = []

    def join(self) -> None:
        torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()


SIDE: Optional[SideStream] = None  # set by engine.TrainStep around backward


def wgrad(dY: Tensor, X: Tensor, out: Tensor, engine: int = ENGINE_AUTO, bias_grad: Optional[Tensor] = None) -> Tensor:
    side = SIDE
    if side is not None and dY.is_cuda and not PROFILER.enabled:
        ready = torch.cuda.Event()
        ready.record()
        side.keep.append((dY, X, out, bias_grad))
        with torch.cuda.stream(side.stream):
            side.stream.wait_event(ready)
            return _wgrad(dY, X, out, engine, bias_grad)
    return _wgrad(dY, X, out, engine, bias_grad)


def _wgrad(dY: Tensor, X: Tensor, out: Tensor, engine: int = ENGINE_AUTO, bias_grad: Optional[Tensor] = None) -> Tensor:
    """out[n,k] += sum_m dY[m,n] * X[m,k]   (out fp32, pre-zeroed by the caller).
    bias_grad (fp32 [n]): += sum_m dY[m,n] in the same launch (one extra N=16 MMA per K step against a tile of ones)."""
    M, N = dY.shape
    K = X.shape[1]
    # output tiles as the tcgen05 kernel cuts them (128 x 64|128|256); one split-K work item per CTA at most: every
    # work item ends with a full tile of fp32 reductions into `out`, so more items than CTAs only adds atomics
    bn = 64 if K <= 64 else (128 if K <= 128 else 256)
    tiles = max(1, math.ceil(N / 128) * math.ceil(K / bn))
    split = max(1, min(math.ceil(M / 256), sm_count() // tiles))
    return gemm(dY.t(), X.t(), out, accumulate=True, split_k=split, engine=engine, row_sum=bias_grad)


_SM = None


def sm_count() -> int:
    global _SM
    if _SM is None:
        _SM = int(_lib.lib().ogv_sm_count())
    return _SM


# --------------------------------------------------------------------------------- layout / misc
def nchw_to_rows(x: Tensor) -> Tensor:
    """[B,C,H,W] (NCHW-contiguous) -> [B*H*W, C] rows."""
    _require_cuda(x)
    B, C, H, W = x.shape
    y = torch.empty((B * H * W, C), device=x.device, dtype=x.dtype)
    _call("ogv_nchw_to_nhwc", _p(x), _p(y), B, C, H * W, dtype_code(x), _stream())
    return y


def rows_to_nchw(y: Tensor, B: int, C: int, H: int, W: int) -> Tensor:
    _require_cuda(y)
    x = torch.empty((B, C, H, W), device=y.device, dtype=y.dtype)
    _call("ogv_nhwc_to_nchw", _p(y), _p(x), B, C, H * W, dtype_code(y), _stream())
    return x


def cast_transpose(src: Tensor, dst: Optional[Tensor], dst_t: Optional[Tensor]) -> None:
    """src fp32 [rows, cols] -> dst[rows, cols] and/or dst_t[cols, rows] (compute dtype; strided views ok)."""
    _require_cuda(src, dst, dst_t)
    _f32(src, "src")
    rows, cols = src.shape
    ref = dst if dst is not None else dst_t
    _call("ogv_cast_transpose", _p(src), _p(dst), dst.stride(0) if dst is not None else 0, _p(dst_t),
                                        dst_t.stride(0) if dst_t is not None else 0, rows, cols, dtype_code(ref),
                                        _stream())


class CastBatch:
    """Device-resident descriptor table for ogv_cast_batch: every (src fp32 [rows, cols] contiguous, dst, dst_t) cast of
    a model re-run by ONE launch per step (dst / dst_t: compute-dtype or fp32 2-D views, either may be None)."""

    def __init__(self, triples):
        import numpy as np
        rec = np.dtype([("src", "<u8"), ("dst", "<u8"), ("dst_t", "<u8"), ("ld_dst", "<i8"), ("ld_dst_t", "<i8"),
                        ("rows", "<i4"), ("cols", "<i4"), ("tile0", "<i4"), ("dtype", "<i4")])
        assert rec.itemsize == 56  # sizeof(ogv_cast_item)
        arr = np.zeros(len(triples), dtype=rec)
        tile0 = 0
        for i, (src, dst, dst_t) in enumerate(triples):
            _require_cuda(src, dst, dst_t)
            _f32(src, "src")
            if src.dim() != 2 or not src.is_contiguous():
                raise ValueError("cast_batch: src must be a contiguous 2-D fp32 tensor")
            ref = dst if dst is not None else dst_t
            for t in (dst, dst_t):
                if t is not None and (t.dim() != 2 or t.stride(1) != 1 or t.dtype != ref.dtype):
                    raise ValueError("cast_batch: dst / dst_t must be 2-D row-major views of one dtype")
            rows, cols = src.shape
            arr[i] = (src.data_ptr(), dst.data_ptr() if dst is not None else 0,
                      dst_t.data_ptr() if dst_t is not None else 0, dst.stride(0) if dst is not None else 0,
                      dst_t.stride(0) if dst_t is not None else 0, rows, cols, tile0, dtype_code(ref))
            tile0 += ((rows + 31) // 32) * ((cols + 31) // 32)
        self.n, self.total_tiles = len(triples), tile0
        self.keep = list(triples)  # the table holds raw pointers: keep the tensors alive
        dev = triples[0][0].device if triples else "cpu"
        self.items = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)

    def run(self) -> None:
        if self.n:
            _call("ogv_cast_batch", ctypes.c_void_p(self.items.data_ptr()), self.n, self.total_tiles, _stream())


def cast(src: Tensor, dtype: torch.dtype) -> Tensor:
    """fp32 [rows, cols] -> compute-dtype copy (ogv_cast_transpose without the transposed output)."""
    dst = torch.empty(src.shape, device=src.device, dtype=dtype)
    cast_transpose(src.contiguous(), dst, None)
    return dst


def rowscale(x: Tensor, scale: Tensor, rows_per_scale: int) -> Tensor:
    _require_cuda(x, scale)
    _f32(scale, "scale")
    y = torch.empty_like(x)
    _call("ogv_rowscale", _p(x), _p(scale), _p(y), x.shape[0], x.shape[1], rows_per_scale, dtype_code(x),
                                  _stream())
    return y


def rowscale_colsum(x: Tensor, scale: Tensor, rows_per_scale: int, out: Tensor, store: bool = True) -> Optional[Tensor]:
    """-> y = x * scale[row // rows_per_scale];  out[n] += sum_m y[m,n]  (one pass).  store=False: reduction only."""
    _require_cuda(x, scale, out)
    _f32(scale, "scale")
    _f32(out, "out")
    y = torch.empty_like(x) if store else None
    _call("ogv_rowscale_colsum", _p(x), _p(scale), _p(y), _p(out), x.shape[0], x.shape[1], rows_per_scale,
          dtype_code(x), _stream())
    return y


def colsum(x: Tensor, out: Tensor) -> Tensor:
    """out[n] += sum_m x[m,n]"""
    _require_cuda(x, out)
    _rows(x, "x")
    _f32(out, "out")
    _call("ogv_colsum", _p(x), x.stride(0), _p(out), x.shape[0], x.shape[1], dtype_code(x), _stream())
    return out


def mul_dact(a: Tensor, pre: Tensor, act: str) -> Tensor:
    out = torch.empty_like(a)
    _call("ogv_mul_dact", _p(a), _p(pre), _p(out), a.numel(), ACT[act], dtype_code(a), _stream())
    return out


def add(a: Tensor, b: Tensor) -> Tensor:
    y = torch.empty_like(a)
    _call("ogv_add", _p(a), _p(b), _p(y), a.numel(), dtype_code(a), _stream())
    return y


# ------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float):
    _require_cuda(x, gamma, beta)
    M, C = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _call("ogv_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), M, C, float(eps),
                                       dtype_code(x), _stream())
    return y, mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dres: Optional[Tensor],
                  dgamma: Tensor, dbeta: Tensor) -> Tensor:
    M, C = x.shape
    dx = torch.empty_like(x)
    _call("ogv_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dgamma),
                                       _p(dbeta), M, C, dtype_code(x), _stream())
    return dx


# --------------------------------------------------------------------------------------- outlook
def outlook_core_fwd(va: Tensor, B: int, H: int, W: int, C: int, heads: int) -> Tensor:
    _require_cuda(va)
    y = torch.empty((va.shape[0], C), device=va.device, dtype=va.dtype)
    _call("ogv_outlook_core_fwd", _p(va), va.stride(0), _p(y), B, H, W, C, heads, dtype_code(va), _stream())
    return y


def outlook_core_bwd(va: Tensor, dy: Tensor, B: int, H: int, W: int, C: int, heads: int) -> Tensor:
    dva = torch.empty_like(va)
    _call("ogv_outlook_core_bwd", _p(va), va.stride(0), _p(dy), _p(dva), B, H, W, C, heads, dtype_code(va),
                                          _stream())
    return dva


# ------------------------------------------------------------------------------------- BatchNorm
def colstats(x: Tensor, ssum: Tensor, ssq: Tensor) -> None:
    _rows(x, "x")
    _call("ogv_colstats", _p(x), x.stride(0), _p(ssum), _p(ssq), x.shape[0], x.shape[1], dtype_code(x),
                                  _stream())


def bn_finalize(ssum, ssq, gamma, beta, running_mean, running_var, scale, shift, mean, rstd, n: int, eps: float,
                momentum: float, training: bool) -> None:
    C = scale.numel()
    _call("ogv_bn_finalize", _p(ssum), _p(ssq), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                     _p(scale), _p(shift), _p(mean), _p(rstd), n, C, float(eps), float(momentum),
                                     int(training), _stream())


def bn_apply(x: Tensor, scale: Tensor, shift: Tensor, res: Optional[Tensor]) -> Tensor:
    out = torch.empty_like(x)
    _call("ogv_bn_apply", _p(x), _p(scale), _p(shift), _p(res), _p(out), x.shape[0], x.shape[1],
                                  dtype_code(x), _stream())
    return out


def bn_bwd_reduce(dy, x, mean, rstd, dgamma, dbeta) -> None:
    _call("ogv_bn_bwd_reduce", _p(dy), _p(x), _p(mean), _p(rstd), _p(dgamma), _p(dbeta), x.shape[0],
                                       x.shape[1], dtype_code(x), _stream())


def bn_bwd_apply(dy, x, mean, rstd, gamma, dgamma, dbeta) -> Tensor:
    dx = torch.empty_like(x)
    _call("ogv_bn_bwd_apply", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dgamma), _p(dbeta), _p(dx),
                                      x.shape[0], x.shape[1], dtype_code(x), _stream())
    return dx


def im2col3x3(x: Tensor, kpad: int) -> Tensor:
    """channels_last [B, Cin, H, W] -> 3x3 / stride 1 / pad 1 patches as rows [B*H*W, kpad] (column (ky*3+kx)*Cin + ci)."""
    _require_cuda(x)
    B, Cin, H, W = x.shape
    if not x.permute(0, 2, 3, 1).is_contiguous():
        raise ValueError("im2col3x3 needs a channels_last tensor")
    cols = torch.empty((B * H * W, kpad), device=x.device, dtype=x.dtype)
    _call("ogv_im2col3x3", _p(x), _p(cols), B, H, W, Cin, kpad, dtype_code(x), _stream())
    return cols


def im2col3x3_vec(x: Tensor, stride: int) -> Tensor:
    """channels_last [B, Cin, H, W] (Cin % 8 == 0) -> 3x3 / pad 1 / stride 1|2 patches as rows [B*Ho*Wo, 9*Cin]."""
    _require_cuda(x)
    B, Cin, H, W = x.shape
    if not x.permute(0, 2, 3, 1).is_contiguous():
        raise ValueError("im2col3x3_vec needs a channels_last tensor")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    cols = torch.empty((B * Ho * Wo, 9 * Cin), device=x.device, dtype=x.dtype)
    _call("ogv_im2col3x3_vec", _p(x), _p(cols), B, H, W, Cin, stride, dtype_code(x), _stream())
    return cols


def col2im3x3_vec(dcols: Tensor, B: int, H: int, W: int, Cin: int, stride: int) -> Tensor:
    """dcols [B*Ho*Wo, 9*Cin] -> input gradient as rows [B*H*W, Cin] (the transpose of im2col3x3_vec)."""
    _require_cuda(dcols)
    _rows(dcols, "dcols")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if tuple(dcols.shape) != (B * Ho * Wo, 9 * Cin) or not dcols.is_contiguous():
        raise ValueError(f"col2im3x3_vec: dcols must be a contiguous [{B * Ho * Wo}, {9 * Cin}] tensor")
    dx = torch.empty((B * H * W, Cin), device=dcols.device, dtype=dcols.dtype)
    _call("ogv_col2im3x3_vec", _p(dcols), _p(dx), B, H, W, Cin, stride, dtype_code(dcols), _stream())
    return dx


def conv3x3_supported(x: Tensor, Co: int, stride: int) -> bool:
    """True when the implicit-GEMM convolution (ogv_conv3x3_fwd / _wgrad) serves this channels_last bf16 image."""
    if x.dtype != torch.bfloat16 or os.environ.get("OGV_CONV_IMPLICIT", "1") == "0":
        return False
    B, Cin, H, W = x.shape
    return bool(_lib.lib().ogv_conv3x3_supported(B, H, W, Cin, Co, stride))


def conv3x3_fwd(x: Tensor, w2: Tensor, stride: int, col_sum: Optional[Tensor] = None,
                col_sumsq: Optional[Tensor] = None) -> Tensor:
    """channels_last bf16 x [B, Cin, H, W], w2 [Co, 9*Cin] (column (ky*3+kx)*Cin + c) -> rows [B*Ho*Wo, Co]: the 3x3 / pad 1
    convolution as an implicit GEMM (the patch matrix is never written; 5-D TMA boxes of x feed the tcgen05 pipeline)."""
    _require_cuda(x, w2)
    B, Cin, H, W = x.shape
    Co = w2.shape[0]
    if not x.permute(0, 2, 3, 1).is_contiguous() or not w2.is_contiguous() or w2.shape[1] != 9 * Cin:
        raise ValueError("conv3x3_fwd needs a channels_last image and a contiguous [Co, 9*Cin] weight")
    Mo = B * (H // stride) * (W // stride)
    y = torch.empty((Mo, Co), device=x.device, dtype=x.dtype)
    if PROFILER.enabled:
        PROFILER.cur_flops = 2 * Mo * Co * 9 * Cin
        PROFILER.cur_label = f"ogv_conv3x3_fwd[{Mo}x{Co}x{9 * Cin} s{stride}]"
        PROFILER.cur_kernel = f"gemm_tc_kernel<{64 if Co <= 64 else (128 if Co <= 128 else 256)}, __nv_bfloat16>"
    _f32(col_sum, "col_sum")
    _f32(col_sumsq, "col_sumsq")
    _call("ogv_conv3x3_fwd", _p(x), _p(w2), _p(y), _p(col_sum), _p(col_sumsq), B, H, W, Cin, Co, stride, _stream())
    return y


def conv3x3_wgrad(x: Tensor, dy: Tensor, dw2: Tensor, stride: int) -> Tensor:
    """dw2[co, (ky*3+kx)*Cin + c] += sum over output pixels of dy[pixel, co] * patch[pixel, ...]  (dw2 fp32, pre-zeroed)."""
    _require_cuda(x, dy, dw2)
    B, Cin, H, W = x.shape
    Mo, Co = dy.shape
    _f32(dw2, "dw2")
    if (not x.permute(0, 2, 3, 1).is_contiguous() or not dy.is_contiguous() or tuple(dw2.shape) != (Co, 9 * Cin)
            or Mo != B * (H // stride) * (W // stride)):
        raise ValueError("conv3x3_wgrad: shape / layout mismatch")
    if PROFILER.enabled:
        PROFILER.cur_flops = 2 * Mo * Co * 9 * Cin
        PROFILER.cur_label = f"ogv_conv3x3_wgrad[{Co}x{9 * Cin}x{Mo} s{stride}]"
        PROFILER.cur_kernel = "gemm_tc_kernel<256, float>"
    _call("ogv_conv3x3_wgrad", _p(x), _p(dy), _p(dw2), B, H, W, Cin, Co, stride, _stream())
    return dw2


def bn_act_apply(x: Tensor, scale: Tensor, shift: Tensor, act: str) -> Tensor:
    """out = act(scale*x + shift) on rows [M, C]  (the BN + act of stem_head.py:23-32 / downsampling.py:28-65)"""
    _rows(x, "x")
    out = torch.empty_like(x)
    _call("ogv_bn_act_apply", _p(x), _p(scale), _p(shift), _p(out), x.shape[0], x.shape[1], ACT[act], dtype_code(x),
          _stream())
    return out


def bn_act_bwd_reduce(dy, x, scale, shift, mean, rstd, dgamma, dbeta, act: str) -> None:
    _call("ogv_bn_act_bwd_reduce", _p(dy), _p(x), _p(scale), _p(shift), _p(mean), _p(rstd), _p(dgamma), _p(dbeta),
          x.shape[0], x.shape[1], ACT[act], dtype_code(x), _stream())


def bn_act_bwd_apply(dy, x, scale, shift, mean, rstd, gamma, dgamma, dbeta, act: str) -> Tensor:
    dx = torch.empty_like(x)
    _call("ogv_bn_act_bwd_apply", _p(dy), _p(x), _p(scale), _p(shift), _p(mean), _p(rstd), _p(gamma), _p(dgamma),
          _p(dbeta), _p(dx), x.shape[0], x.shape[1], ACT[act], dtype_code(x), _stream())
    return dx


# ---------------------------------------------------------------------------------------- MBConv
def dwconv_fwd(e_pre, scale1, shift1, w, ssum2, ssq2, B, H, W, act: str) -> Tensor:
    d_pre = torch.empty_like(e_pre)
    _call("ogv_dwconv_fwd", _p(e_pre), _p(scale1), _p(shift1), _p(w), _p(d_pre), _p(ssum2), _p(ssq2), B, H, W,
                                    e_pre.shape[1], ACT[act], dtype_code(e_pre), _stream())
    return d_pre


def dwconv_bwd(dd_pre, e_pre, scale1, shift1, mean1, rstd1, w, dw, dgamma1, dbeta1, B, H, W, act: str) -> Tensor:
    du1 = torch.empty_like(e_pre)
    _call("ogv_dwconv_bwd", _p(dd_pre), _p(e_pre), _p(scale1), _p(shift1), _p(mean1), _p(rstd1), _p(w),
                                    _p(du1), _p(dw), _p(dgamma1), _p(dbeta1), B, H, W, e_pre.shape[1], ACT[act],
                                    dtype_code(e_pre), _stream())
    return du1


def se_pool(d_pre, scale2, shift2, B, HW, act: str) -> Tensor:
    pool = torch.empty((B, d_pre.shape[1]), device=d_pre.device, dtype=torch.float32)
    _call("ogv_se_pool", _p(d_pre), _p(scale2), _p(shift2), _p(pool), B, HW, d_pre.shape[1], ACT[act],
                                 dtype_code(d_pre), _stream())
    return pool


def se_mlp_supported(Cm: int, Cs: int, dtype: torch.dtype) -> bool:
    """True when the one-kernel squeeze-excite MLP serves this shape (bf16 compute mode, small weight matrices).
    Measured in the replayed step of cfg 2: the FFMA kernel wins at Cm <= 512 (stages 0-1: 23-27 us for what were two
    tensor-core GEMM launches, two casts and two activations); at Cm = 1024 it takes 66 us and the GEMM route is as fast.
    OGV_SE_FUSED_MAXCM moves the cut, OGV_SE_FUSED=0 disables the kernel."""
    if os.environ.get("OGV_SE_FUSED", "1") == "0" or Cm > int(os.environ.get("OGV_SE_FUSED_MAXCM", "512")):
        return False
    return bool(_lib.lib().ogv_se_mlp_supported(Cm, Cs, BF16 if dtype == torch.bfloat16 else F32))


def se_mlp_fwd(pool: Tensor, w1t: Tensor, b1: Tensor, w2t: Tensor, b2: Tensor, act: str):
    """pool [B, Cm] fp32; w1t = W1^T [Cm, Cs], w2t = W2^T [Cs, Cm] bf16 (contiguous).
    -> pool_c, s1_pre, s1a (bf16), gate_pre, gate (fp32)"""
    B, Cm = pool.shape
    Cs = w1t.shape[1]
    dev = pool.device
    pool_c = torch.empty((B, Cm), device=dev, dtype=torch.bfloat16)
    s1_pre = torch.empty((B, Cs), device=dev, dtype=torch.bfloat16)
    s1a = torch.empty((B, Cs), device=dev, dtype=torch.bfloat16)
    gate_pre = torch.empty((B, Cm), device=dev, dtype=torch.float32)
    gate = torch.empty((B, Cm), device=dev, dtype=torch.float32)
    _call("ogv_se_mlp_fwd", _p(pool), _p(w1t), _p(b1), _p(w2t), _p(b2), _p(pool_c), _p(s1_pre), _p(s1a), _p(gate_pre),
          _p(gate), B, Cm, Cs, ACT[act], _stream())
    return pool_c, s1_pre, s1a, gate_pre, gate


def se_mlp_bwd(dgate: Tensor, gate_pre: Tensor, s1_pre: Tensor, w2: Tensor, w1: Tensor, act: str):
    """dgate, gate_pre [B, Cm] fp32; s1_pre [B, Cs] bf16; w2 = W2 [Cm, Cs], w1 = W1 [Cs, Cm] bf16 (contiguous).
    -> dgate_c [B, Cm] bf16, ds1_pre [B, Cs] bf16, dpool [B, Cm] fp32"""
    B, Cm = dgate.shape
    Cs = s1_pre.shape[1]
    dev = dgate.device
    dgate_c = torch.empty((B, Cm), device=dev, dtype=torch.bfloat16)
    ds1_pre = torch.empty((B, Cs), device=dev, dtype=torch.bfloat16)
    dpool = torch.empty((B, Cm), device=dev, dtype=torch.float32)
    _call("ogv_se_mlp_bwd", _p(dgate), _p(gate_pre), _p(s1_pre), _p(w2), _p(w1), _p(dgate_c), _p(ds1_pre), _p(dpool), B,
          Cm, Cs, ACT[act], _stream())
    return dgate_c, ds1_pre, dpool


def bn_act_gate(d_pre, scale2, shift2, gate, B, HW, act: str) -> Tensor:
    d_act = torch.empty_like(d_pre)
    _call("ogv_bn_act_gate", _p(d_pre), _p(scale2), _p(shift2), _p(gate), _p(d_act), B, HW, d_pre.shape[1],
                                     ACT[act], dtype_code(d_pre), _stream())
    return d_act


def se_bwd_reduce(dd_act, d_pre, scale2, shift2, B, HW, act: str) -> Tensor:
    dgate = torch.empty((B, d_pre.shape[1]), device=d_pre.device, dtype=torch.float32)
    _call("ogv_se_bwd_reduce", _p(dd_act), _p(d_pre), _p(scale2), _p(shift2), _p(dgate), B, HW,
                                       d_pre.shape[1], ACT[act], dtype_code(d_pre), _stream())
    return dgate


def mbconv_bwd_stats(dd_act, d_pre, scale2, shift2, mean2, rstd2, B, HW, act: str) -> Tensor:
    """-> stats [5, B, Cm] fp32 (stats[0] = dgate); see include/ogv.h."""
    stats = torch.empty((5, B, d_pre.shape[1]), device=d_pre.device, dtype=torch.float32)
    _call("ogv_mbconv_bwd_stats", _p(dd_act), _p(d_pre), _p(scale2), _p(shift2), _p(mean2), _p(rstd2), _p(stats), B,
          HW, d_pre.shape[1], ACT[act], dtype_code(d_pre), _stream())
    return stats


def mbconv_bn2_finalize(stats, gate, dpool, dgamma2, dbeta2, B, HW) -> None:
    _call("ogv_mbconv_bn2_finalize", _p(stats), _p(gate), _p(dpool), _p(dgamma2), _p(dbeta2), B, HW, gate.shape[1],
          _stream())


def dw_bn2_bwd_apply(dd_act, d_pre, gate, dpool, scale2, shift2, mean2, rstd2, gamma2, dgamma2, dbeta2, B, HW,
                     act: str) -> Tensor:
    dd_pre = torch.empty_like(d_pre)
    _call("ogv_dw_bn2_bwd_apply", _p(dd_act), _p(d_pre), _p(gate), _p(dpool), _p(scale2), _p(shift2), _p(mean2),
          _p(rstd2), _p(gamma2), _p(dgamma2), _p(dbeta2), _p(dd_pre), B, HW, d_pre.shape[1], ACT[act],
          dtype_code(d_pre), _stream())
    return dd_pre


# --------------------------------------------------------------------------------- grid attention
def grid_attn_fwd(qkv: Tensor, B, H, W, C, heads, g) -> Tensor:
    _require_cuda(qkv)
    out = torch.empty((qkv.shape[0], C), device=qkv.device, dtype=qkv.dtype)
    _call("ogv_grid_attn_fwd", _p(qkv), _p(out), B, H, W, C, heads, g, dtype_code(qkv), _stream())
    return out


def grid_attn_bwd(qkv: Tensor, dout: Tensor, B, H, W, C, heads, g) -> Tensor:
    dqkv = torch.empty_like(qkv)
    _call("ogv_grid_attn_bwd", _p(qkv), _p(dout), _p(dqkv), B, H, W, C, heads, g, dtype_code(qkv), _stream())
    return dqkv


def grid_attn_probs(qkv: Tensor, B, H, W, C, heads, g) -> Tensor:
    N = (H // g) * (W // g)
    attn = torch.empty((B * g * g, heads, N, N), device=qkv.device, dtype=torch.float32)
    _call("ogv_grid_attn_probs", _p(qkv), _p(attn), B, H, W, C, heads, g, dtype_code(qkv), _stream())
    return attn


# -------------------------------------------------------------------------------------- fused MLP
def mlp_fused_supported(C: int, hidden: int, dtype: torch.dtype) -> bool:
    return dtype == torch.bfloat16 and bool(_lib.lib().ogv_mlp_fused_supported(int(C), int(hidden)))


def mlp_fwd(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, *, act: str, residual: Optional[Tensor] = None,
            row_scale: Optional[Tensor] = None, rows_per_scale: int = 1) -> Tensor:
    """y = residual + row_scale * (act(x w1^T + b1) w2^T + b2) in one kernel; the hidden activation stays on chip.
    x [M, C] bf16 rows, w1 [hidden, C] / w2 [C, hidden] bf16 contiguous, biases fp32."""
    _require_cuda(x, w1, b1, w2, b2, residual, row_scale)
    M, C = x.shape
    Hd = w1.shape[0]
    if x.dtype != torch.bfloat16 or w1.dtype != torch.bfloat16 or w2.dtype != torch.bfloat16:
        raise TypeError("mlp_fwd: bf16 activations and weights only")
    if tuple(w1.shape) != (Hd, C) or tuple(w2.shape) != (C, Hd) or not w1.is_contiguous() or not w2.is_contiguous():
        raise ValueError("mlp_fwd: w1 must be [hidden, C] and w2 [C, hidden], both contiguous")
    _rows(x, "x")
    _f32(b1, "b1")
    _f32(b2, "b2")
    _f32(row_scale, "row_scale")
    if residual is not None:
        _rows(residual, "residual")
        if residual.dtype != x.dtype or tuple(residual.shape) != (M, C):
            raise ValueError("mlp_fwd: residual must match x")
    y = torch.empty((M, C), device=x.device, dtype=x.dtype)
    if PROFILER.enabled:
        PROFILER.cur_bytes = 2 * (M * C * (2 + (residual is not None)) + 2 * Hd * C)
        PROFILER.cur_flops = 4 * M * C * Hd
        PROFILER.cur_kernel = f"mlp_fwd_kernel<{C}>"
    _call("ogv_mlp_fwd", ctypes.c_void_p(x.data_ptr()), x.stride(0), ctypes.c_void_p(w1.data_ptr()), ctypes.c_void_p(b1.data_ptr()),
          ctypes.c_void_p(w2.data_ptr()), ctypes.c_void_p(b2.data_ptr()),
          ctypes.c_void_p(residual.data_ptr()) if residual is not None else None,
          residual.stride(0) if residual is not None else 0,
          ctypes.c_void_p(row_scale.data_ptr()) if row_scale is not None else None, int(rows_per_scale),
          ctypes.c_void_p(y.data_ptr()), y.stride(0), M, C, Hd, ACT[act], _stream())
    return y


def mlp_bwd(xn: Tensor, dy: Tensor, w1: Tensor, w2t: Tensor, w1t: Tensor, b1: Tensor, *, act: str,
            row_scale: Optional[Tensor] = None, rows_per_scale: int = 1):
    """Backward of the fused MLP with the hidden activation RECOMPUTED on chip:
        z = xn w1^T + b1;  dz = (dy w2) * act'(z) * s;  hs = act(z) * s;  dxn = dz w1      (s = row_scale or 1)
    -> (dxn [M, C], dz [M, hidden], hs [M, hidden]); dz / hs feed the weight-gradient GEMMs dW1 = dz^T xn, dW2 = dy^T hs.
    w2t = w2^T [hidden, C], w1t = w1^T [C, hidden] (the pre-transposed copies the dgrad GEMMs use)."""
    _require_cuda(xn, dy, w1, w2t, w1t, b1, row_scale)
    M, C = xn.shape
    Hd = w1.shape[0]
    for name, t in (("xn", xn), ("dy", dy), ("w1", w1), ("w2t", w2t), ("w1t", w1t)):
        if t.dtype != torch.bfloat16:
            raise TypeError(f"mlp_bwd: {name} must be bf16")
    if tuple(w1.shape) != (Hd, C) or tuple(w2t.shape) != (Hd, C) or tuple(w1t.shape) != (C, Hd) or tuple(dy.shape) != (M, C):
        raise ValueError("mlp_bwd: shape mismatch")
    if not (w1.is_contiguous() and w2t.is_contiguous() and w1t.is_contiguous()):
        raise ValueError("mlp_bwd: weights must be contiguous")
    _rows(xn, "xn")
    _rows(dy, "dy")
    _f32(b1, "b1")
    _f32(row_scale, "row_scale")
    dz = torch.empty((M, Hd), device=xn.device, dtype=xn.dtype)
    hs = torch.empty((M, Hd), device=xn.device, dtype=xn.dtype)
    dxn = torch.empty((M, C), device=xn.device, dtype=xn.dtype)
    if PROFILER.enabled:
        PROFILER.cur_bytes = 2 * (3 * M * C + 2 * M * Hd + 3 * Hd * C)
        PROFILER.cur_flops = 6 * M * C * Hd
        PROFILER.cur_kernel = f"mlp_bwd_kernel<{C}>"
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    _call("ogv_mlp_bwd", vp(xn), xn.stride(0), vp(dy), dy.stride(0), vp(w1), vp(w2t), vp(w1t), vp(b1), vp(row_scale),
          int(rows_per_scale), vp(dz), vp(hs), vp(dxn), dxn.stride(0), M, C, Hd, ACT[act], _stream())
    return dxn, dz, hs


# ----------------------------------------------------------------------- flat-arena train-step tail
def sumsq(g: Tensor, out: Tensor) -> None:
    """out[0] += sum g^2 (g: flat fp32 arena)."""
    _require_cuda(g, out)
    _f32(g, "g")
    _f32(out, "out")
    _call("ogv_sumsq", _p(g), g.numel(), _p(out), _stream())


def adamw_flat(p: Tensor, g: Tensor, m: Tensor, v: Tensor, decay_bits: Tensor, hyper: Tensor, gnorm_sq: Optional[Tensor],
               loss: Optional[Tensor], beta1: float, beta2: float, eps: float, weight_decay: float,
               skipped: Optional[Tensor]) -> None:
    """Clip + AdamW over flat fp32 arenas; lr / bias corrections / 1/world / max_norm come from the DEVICE tensor
    `hyper` (see include/ogv.h), so the launch is CUDA-graph replayable under a changing schedule."""
    _require_cuda(p, g, m, v, decay_bits, hyper, gnorm_sq, loss, skipped)
    for name, t in (("p", p), ("g", g), ("m", m), ("v", v), ("hyper", hyper)):
        _f32(t, name)
    if loss is not None and (loss.dtype != torch.float32 or loss.numel() != 1):
        raise TypeError("adamw_flat: loss must be a single float32 value on the device")
    if decay_bits.dtype != torch.int32:
        raise TypeError("adamw_flat: decay_bits must be int32 words")
    _call("ogv_adamw_flat", _p(p), _p(g), _p(m), _p(v), _p(decay_bits), p.numel(), _p(hyper), _p(gnorm_sq), _p(loss),
          float(beta1), float(beta2), float(eps), float(weight_decay), _p(skipped), _stream())


def xent_fwd(logits: Tensor, labels: Tensor, smoothing: float, loss: Tensor) -> Tensor:
    """loss[0] += mean cross entropy with label smoothing of logits [B, K] (fp32 / bf16) vs int64 labels; -> lse [B] fp32."""
    _require_cuda(logits, labels, loss)
    _rows(logits, "logits")
    _f32(loss, "loss")
    if labels.dtype != torch.int64 or labels.dim() != 1 or labels.shape[0] != logits.shape[0] or not labels.is_contiguous():
        raise ValueError("xent_fwd: labels must be a contiguous int64 vector with one entry per row")
    B, K = logits.shape
    lse = torch.empty(B, device=logits.device, dtype=torch.float32)
    _call("ogv_xent_fwd", _p(logits), logits.stride(0), _p(labels), B, K, float(smoothing), dtype_code(logits), _p(lse),
          _p(loss), _stream())
    return lse


def xent_bwd(logits: Tensor, labels: Tensor, lse: Tensor, gout: Optional[Tensor], smoothing: float) -> Tensor:
    """-> dlogits [B, K] (dtype of logits) = gout / B * (softmax - (1-eps)*onehot - eps/K)."""
    _require_cuda(logits, labels, lse, gout)
    _f32(gout, "gout")
    B, K = logits.shape
    dl = torch.empty((B, K), device=logits.device, dtype=logits.dtype)
    _call("ogv_xent_bwd", _p(logits), logits.stride(0), _p(labels), _p(lse), _p(gout), B, K, float(smoothing),
          dtype_code(logits), _p(dl), dl.stride(0), _stream())
    return dl


def train_metrics(logits: Tensor, labels: Tensor, loss: Optional[Tensor], acc: Tensor) -> None:
    """acc[5] += {loss*B, top-1, top-3, top-5 hits, B} (fp32 logits [B, K], int64 labels)."""
    _require_cuda(logits, labels, loss, acc)
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
        raise TypeError("train_metrics: logits must be fp32 [B, K] with unit column stride")
    if labels.dtype != torch.int64 or not labels.is_contiguous():
        raise TypeError("train_metrics: labels must be contiguous int64")
    _f32(acc, "acc")
    _call("ogv_train_metrics", _p(logits), logits.stride(0), _p(labels), logits.shape[0], logits.shape[1], _p(loss),
          _p(acc), _stream())


# ----------------------------------------------------------------------------------------- AdamW
def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step: int, grad_scale: float = 1.0) -> None:
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    _call("ogv_adamw", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                               grad_scale, _stream())
