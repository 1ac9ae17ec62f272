"""Autograd functions for the four fused branches of an OutGridBlock, on "rows" tensors
([B*H*W, C], NHWC order).  Forward and backward are sequences of libogvit kernel launches; the
backward formulas are the ones autograd derives for the reference modules.

    outlook_branch : x + s * proj(outlook_core(softmax(attn(LN x)), v(LN x)))   outlook_attention.py:91-124, Outlook_Block.py:61-63
    mlp_branch     : x + s * fc2(act(fc1(LN x)))                                 outlook_attention.py:43-49, Out_Grid_Block.py:24-32
    mbconv         : x + BN3(Wp (SE(act(BN2(DW(act(BN1(We x))))))))              mbc_conv.py:90-98
    grid_branch    : x + s * proj(grid_mhsa(qkv(LN x)))                          grid_attention.py:62-89,112-131

`s` is the per-sample stochastic-depth scale (mask / keep_prob) or None.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import ops
from ._lib import ENGINE_AUTO, ENGINE_SIMT

Tensor = torch.Tensor


@dataclass
class Geom:
    B: int
    H: int
    W: int

    @property
    def P(self) -> int:
        return self.H * self.W

    @property
    def M(self) -> int:
        return self.B * self.H * self.W


class PreparedLinear:
    """Compute-dtype copies of one [out, in] weight: `w` ([out,in]) and `wt` ([in,out])."""

    __slots__ = ("w", "wt")

    def __init__(self, w: Tensor, wt: Tensor):
        self.w = w
        self.wt = wt


# fc1 -> act -> fc2 in one kernel with the hidden tile on chip (ogv_mlp_fwd / ogv_mlp_bwd) for the shapes it serves;
# OGV_MLP_FUSED=0 keeps the two-GEMM route (A/B measurements).
import os as _os

MLP_FUSED = _os.environ.get("OGV_MLP_FUSED", "1") != "0"

# When a list, every cast issued by prepare_* below is also recorded as (src, dst, dst_t): modules._Prep keeps the
# records so that modules.refresh_prepared() can re-run all of a model's casts in ONE launch per step.
_CAST_LOG: Optional[list] = None


def _cast_logged(src: Tensor, dst: Optional[Tensor], dst_t: Optional[Tensor]) -> None:
    ops.cast_transpose(src, dst, dst_t)
    if _CAST_LOG is not None:
        _CAST_LOG.append((src, dst, dst_t))


def prepare_linear(weight: Tensor, dtype: torch.dtype, out_rows: Optional[int] = None) -> PreparedLinear:
    """weight: fp32 parameter [out, in(,1,1)].  For fp32 compute the parameter itself is used."""
    w2 = weight.detach().reshape(weight.shape[0], -1)
    n, k = w2.shape
    rows = out_rows or n
    if dtype == torch.float32 and rows == n:
        return PreparedLinear(w2, w2.t())
    alloc = torch.empty if rows == n else torch.zeros  # zero fill only where padding rows exist
    w = alloc((rows, k), device=w2.device, dtype=dtype)
    wt = alloc((k, rows), device=w2.device, dtype=dtype)
    _cast_logged(w2.contiguous(), w[:n], wt[:, :n])
    return PreparedLinear(w, wt)


def _scaled_grad(dy: Tensor, scale: Optional[Tensor], rows_per_sample: int, dbias: Tensor) -> Tensor:
    """gy = dy * per-sample DropPath scale, and dbias += colsum(gy) -- one pass when a scale is present."""
    if scale is None:
        ops.colsum(dy, dbias)
        return dy
    return ops.rowscale_colsum(dy, scale, rows_per_sample, dbias)


def _zeros(shape, like: Tensor) -> Tensor:
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


# Engine hooks (engine.FlatState / ddp.ArenaGradAllReduce).  With a module in "direct" mode every parameter's .grad is
# a view of ONE flat fp32 arena that the train step zeroes with a single memset: the backward kernels (all of them
# accumulate: split-K `red`, column-sum atomics) add straight into those views and the autograd functions return None
# for the parameters -- no per-branch zero-filled arenas (~190 fill launches per step), no AccumulateGrad copies, no
# pack / unpack around the gradient all-reduce.
SCRATCH = None     # object with .take(n) -> zeroed fp32 view or None (pre-zeroed by the step's memset)
GRAD_READY = None  # callable(list of parameters whose gradients are complete) -- the all-reduce bucket trigger


def _scratch_zeros(n: int, like: Tensor) -> Tensor:
    if SCRATCH is not None:
        t = SCRATCH.take(n)
        if t is not None:
            return t
    return _zeros(n, like)


def _notify(params) -> None:
    if GRAD_READY is not None:
        GRAD_READY([p for p in params if p is not None])


def _is_direct(meta, params) -> bool:
    return bool(meta.get("direct")) and all(p is None or p.grad is not None for p in params)


def _empty(shape, like: Tensor, dtype=None) -> Tensor:
    return torch.empty(shape, device=like.device, dtype=dtype or like.dtype)


# =================================================================================================
# MLP branch (MLP2d of the Outlooker and the BHWC MLP share the same math on rows)
# =================================================================================================
class MlpBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, scale, meta):
        p1: PreparedLinear = meta["p1"]
        p2: PreparedLinear = meta["p2"]
        act, P = meta["act"], meta["rows_per_sample"]
        with_ln, with_res = ln_w is not None, meta["with_res"]
        x = x.contiguous()
        M, C = x.shape
        Hd = p1.w.shape[0]
        ops.PROFILER.tag = ("K4a" if Hd <= 2 * C else "K4b", "fwd", M, C)
        if with_ln:
            xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta["eps"])
        else:
            xn, mean, rstd = x, None, None
        fused = MLP_FUSED and ops.mlp_fused_supported(C, Hd, x.dtype) and b1 is not None and b2 is not None
        if fused:
            # fc1 -> act -> fc2 -> DropPath scale -> residual in ONE kernel; the [M, hidden] activation never leaves the
            # SM, and nothing wide is saved: the backward kernel recomputes it from the (narrow) LayerNorm output
            y = ops.mlp_fwd(xn, p1.w, b1, p2.w, b2, act=act, residual=x if with_res else None, row_scale=scale,
                            rows_per_scale=P)
            z = h = None
        else:
            # the forward epilogue saves act'(z) (it has the transcendental in hand), so the backward epilogue is a
            # plain multiply instead of a second erf/exp evaluation per element
            z = _empty((M, Hd), x)
            h = _empty((M, Hd), x)
            ops.gemm(xn, p1.w, h, bias=b1, pre_out=z, act=act, pre_out_grad=True)
            y = _empty((M, C), x)
            ops.gemm(h, p2.w, y, bias=b2, row_scale=scale, rows_per_scale=P, residual=x if with_res else None)
        ctx.fused = fused
        ctx.meta = meta
        ctx.params = (ln_w, ln_b, w1, b1, w2, b2)
        ctx.save_for_backward(x, xn, mean, rstd, z, h, ln_w, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        meta = ctx.meta
        x, xn, mean, rstd, z, h, ln_w, scale = ctx.saved_tensors
        p1: PreparedLinear = meta["p1"]
        p2: PreparedLinear = meta["p2"]
        act, P = meta["act"], meta["rows_per_sample"]
        with_ln, with_res = ln_w is not None, meta["with_res"]
        dy = dy.contiguous()
        M, C = x.shape
        Hd = p1.w.shape[0]
        ops.PROFILER.tag = ("K4a" if Hd <= 2 * C else "K4b", "bwd", M, C)
        direct = _is_direct(meta, ctx.params)
        if direct:  # accumulate straight into the parameters' views of the flat gradient arena
            pl, pb, pw1, pb1, pw2, pb2 = ctx.params
            dW1, dW2, db1, db2 = pw1.grad.view(Hd, C), pw2.grad.view(C, Hd), pb1.grad, pb2.grad
            dg, dbt = (pl.grad, pb.grad) if with_ln else (None, None)
        else:  # one zeroed fp32 arena for all parameter gradients of the branch
            arena = _zeros(Hd * C * 2 + Hd + C + 2 * C, x)
            o = 0
            dW1 = arena[o:o + Hd * C].view(Hd, C); o += Hd * C
            dW2 = arena[o:o + C * Hd].view(C, Hd); o += C * Hd
            db1 = arena[o:o + Hd]; o += Hd
            db2 = arena[o:o + C]; o += C
            dg = arena[o:o + C]; o += C
            dbt = arena[o:o + C]; o += C
        if ctx.fused and (with_ln or not with_res):
            # one kernel: recompute z, h on chip; dz = (dy W2) * act'(z) * s, hs = act(z) * s, dxn = dz W1.  The DropPath
            # scale s rides in dz / hs, so dy itself is the operand of the fc2 weight gradient (no scaled copy of dy)
            if scale is None:
                ops.colsum(dy, db2)
            else:
                ops.rowscale_colsum(dy, scale, P, db2, store=False)
            dxn, dz, hs = ops.mlp_bwd(xn, dy, p1.w, p2.wt, p1.wt, ctx.params[3].detach(), act=act, row_scale=scale,
                                      rows_per_scale=P)
            ops.wgrad(dy, hs, dW2)
            ops.wgrad(dz, xn, dW1, bias_grad=db1)  # db1 = colsum(dz) rides with the weight gradient
        else:
            if ctx.fused:  # shapes the fused backward does not serve: recompute the saved pair the unfused way
                z = _empty((M, Hd), x)
                h = _empty((M, Hd), x)
                ops.gemm(xn, p1.w, h, bias=ctx.params[3].detach(), pre_out=z, act=act, pre_out_grad=True)
            gy = _scaled_grad(dy, scale, P, db2)
            # fc2 backward (+ activation derivative and the fc1 bias gradient fused into the dgrad epilogue)
            dz = _empty((M, Hd), x)
            # bias gradient as a separate full-rate pass: the column sum fused into this wide-N epilogue costs about
            # twice what the streaming reduction does (measured 165 us vs 81 us at stage 0)
            ops.gemm(gy, p2.wt, dz, dact_src=z, dact="mul")
            ops.wgrad(gy, h, dW2)
            # fc1 backward
            dxn = _empty((M, C), x)
            if with_ln:
                ops.gemm(dz, p1.wt, dxn)
            else:
                ops.gemm(dz, p1.wt, dxn, residual=dy if with_res else None)
            ops.wgrad(dz, xn, dW1, bias_grad=db1)
        dx = ops.layernorm_bwd(dxn, x, ln_w, mean, rstd, dy if with_res else None, dg, dbt) if with_ln else dxn
        if direct:
            _notify(ctx.params)
            return (dx,) + (None,) * 8
        if with_ln:
            return dx, dg, dbt, dW1.view(meta["w1_shape"]), db1, dW2.view(meta["w2_shape"]), db2, None, None
        return dxn, None, None, dW1.view(meta["w1_shape"]), db1, dW2.view(meta["w2_shape"]), db2, None, None


def mlp_branch(x: Tensor, ln_w, ln_b, w1, b1, w2, b2, scale, *, p1, p2, eps: float, act: str, rows_per_sample: int,
               with_res: bool, direct: bool = False) -> Tensor:
    meta = dict(p1=p1, p2=p2, eps=eps, act=act, rows_per_sample=rows_per_sample, with_res=with_res, direct=direct,
                w1_shape=tuple(w1.shape), w2_shape=tuple(w2.shape))
    return MlpBranchFn.apply(x, ln_w, ln_b, w1, b1, w2, b2, scale, meta)


# =================================================================================================
# Outlook attention branch
# =================================================================================================
def outlook_npad(C: int, heads: int) -> int:
    return (C + 9 * heads + 7) // 8 * 8


class OutlookBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, wv, bv, wa, ba, wp, bp, scale, meta):
        pva: PreparedLinear = meta["pva"]   # rows: [v (C) | attn (9*heads) | zero pad]
        pp: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        heads, with_res = meta["heads"], meta["with_res"]
        with_ln = ln_w is not None
        x = x.contiguous()
        M, C = x.shape
        npad = pva.w.shape[0]
        ops.PROFILER.tag = ("K1", "fwd", M, C, heads)
        if with_ln:
            xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta["eps"])
        else:
            xn, mean, rstd = x, None, None
        va = _empty((M, npad), x)
        ops.gemm(xn, pva.w, va, bias=meta["bva"])
        yc = ops.outlook_core_fwd(va, g.B, g.H, g.W, C, heads)
        y = _empty((M, C), x)
        ops.gemm(yc, pp.w, y, bias=bp, row_scale=scale, rows_per_scale=g.P, residual=x if with_res else None)
        ctx.meta = meta
        ctx.params = (ln_w, ln_b, wv, bv, wa, ba, wp, bp)
        ctx.save_for_backward(x, xn, mean, rstd, va, yc, ln_w, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        meta = ctx.meta
        x, xn, mean, rstd, va, yc, ln_w, scale = ctx.saved_tensors
        pva: PreparedLinear = meta["pva"]
        pp: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        heads, with_res = meta["heads"], meta["with_res"]
        with_ln = ln_w is not None
        dy = dy.contiguous()
        M, C = x.shape
        npad = pva.w.shape[0]
        ops.PROFILER.tag = ("K1", "bwd", M, C, heads)
        nl = 9 * heads
        flat = meta.get("flat")  # (dWva [npad, C], dbva [npad]) views over (v | attn | pad) in the flat gradient arena
        direct = _is_direct(meta, ctx.params) and flat is not None
        if direct:
            pl, pb, pwv, pbv, pwa, pba, pwp, pbp = ctx.params
            dWva, dbva = flat
            dWp, dbp = pwp.grad.view(C, C), pbp.grad
            dg, dbt = (pl.grad, pb.grad) if with_ln else (None, None)
        else:
            arena = _zeros(npad * C + C * C + npad + C + 2 * C, x)
            o = 0
            dWva = arena[o:o + npad * C].view(npad, C); o += npad * C
            dWp = arena[o:o + C * C].view(C, C); o += C * C
            dbva = arena[o:o + npad]; o += npad
            dbp = arena[o:o + C]; o += C
            dg = arena[o:o + C]; o += C
            dbt = arena[o:o + C]; o += C
        gy = _scaled_grad(dy, scale, g.P, dbp)
        dyc = _empty((M, C), x)
        ops.gemm(gy, pp.wt, dyc)
        ops.wgrad(gy, yc, dWp)
        dva = ops.outlook_core_bwd(va, dyc, g.B, g.H, g.W, C, heads)
        dxn = _empty((M, C), x)
        if with_ln:
            ops.gemm(dva, pva.wt, dxn)
        else:
            ops.gemm(dva, pva.wt, dxn, residual=dy if with_res else None)
        ops.wgrad(dva, xn, dWva, bias_grad=dbva)
        if direct:
            dx = ops.layernorm_bwd(dxn, x, ln_w, mean, rstd, dy if with_res else None, dg, dbt) if with_ln else dxn
            _notify(ctx.params)
            return (dx,) + (None,) * 10
        dwv = dWva[:C].reshape(meta["wv_shape"])
        dwa = dWva[C:C + nl].reshape(meta["wa_shape"])
        dbv = dbva[:C] if meta["has_qkv_bias"] else None
        dba = dbva[C:C + nl] if meta["has_qkv_bias"] else None
        if with_ln:
            dx = ops.layernorm_bwd(dxn, x, ln_w, mean, rstd, dy if with_res else None, dg, dbt)
            return dx, dg, dbt, dwv, dbv, dwa, dba, dWp.view(meta["wp_shape"]), dbp, None, None
        return dxn, None, None, dwv, dbv, dwa, dba, dWp.view(meta["wp_shape"]), dbp, None, None


def prepare_outlook_va(wv: Tensor, bv: Optional[Tensor], wa: Tensor, ba: Optional[Tensor], dtype) -> Tuple[PreparedLinear, Tensor]:
    """Concatenate the `v` and `attn` 1x1-conv weights into one zero-padded GEMM operand."""
    C = wv.shape[0]
    nl = wa.shape[0]
    npad = (C + nl + 7) // 8 * 8
    w = torch.zeros((npad, C), device=wv.device, dtype=dtype)
    wt = torch.zeros((C, npad), device=wv.device, dtype=dtype)
    _cast_logged(wv.detach().reshape(C, C).contiguous(), w[:C], wt[:, :C])
    _cast_logged(wa.detach().reshape(nl, C).contiguous(), w[C:C + nl], wt[:, C:C + nl])
    bva = torch.zeros(npad, device=wv.device, dtype=torch.float32)
    if bv is not None:
        bva[:C].copy_(bv.detach())
        if _CAST_LOG is not None:
            _CAST_LOG.append((bv.detach().reshape(1, C), bva[:C].reshape(1, C), None))
    if ba is not None:
        bva[C:C + nl].copy_(ba.detach())
        if _CAST_LOG is not None:
            _CAST_LOG.append((ba.detach().reshape(1, nl), bva[C:C + nl].reshape(1, nl), None))
    return PreparedLinear(w, wt), bva


def outlook_branch(x, ln_w, ln_b, wv, bv, wa, ba, wp, bp, scale, *, pva, bva, pp, geom: Geom, heads: int, eps: float,
                   with_res: bool, direct: bool = False, flat=None) -> Tensor:
    meta = dict(pva=pva, bva=bva, pp=pp, geom=geom, heads=heads, eps=eps, with_res=with_res, direct=direct, flat=flat,
                wv_shape=tuple(wv.shape), wa_shape=tuple(wa.shape), wp_shape=tuple(wp.shape),
                has_qkv_bias=bv is not None)
    return OutlookBranchFn.apply(x, ln_w, ln_b, wv, bv, wa, ba, wp, bp, scale, meta)


# =================================================================================================
# Grid attention branch
# =================================================================================================
class GridBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, wqkv, bqkv, wp, bp, scale, meta):
        pq: PreparedLinear = meta["pq"]
        pp: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        heads, gs, with_res = meta["heads"], meta["grid"], meta["with_res"]
        with_ln = ln_w is not None
        x = x.contiguous()
        M, C = x.shape
        ops.PROFILER.tag = ("K2", "fwd", M, C, (g.H // gs) * (g.W // gs))
        if with_ln:
            xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta["eps"])
        else:
            xn, mean, rstd = x, None, None
        qkv = _empty((M, 3 * C), x)
        ops.gemm(xn, pq.w, qkv, bias=bqkv)
        o = ops.grid_attn_fwd(qkv, g.B, g.H, g.W, C, heads, gs)
        capture = meta.get("capture")
        if capture is not None:
            capture(ops.grid_attn_probs(qkv, g.B, g.H, g.W, C, heads, gs))
        y = _empty((M, C), x)
        ops.gemm(o, pp.w, y, bias=bp, row_scale=scale, rows_per_scale=g.P, residual=x if with_res else None)
        ctx.meta = meta
        ctx.params = (ln_w, ln_b, wqkv, bqkv, wp, bp)
        ctx.save_for_backward(x, xn, mean, rstd, qkv, o, ln_w, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        meta = ctx.meta
        x, xn, mean, rstd, qkv, o, ln_w, scale = ctx.saved_tensors
        pq: PreparedLinear = meta["pq"]
        pp: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        heads, gs, with_res = meta["heads"], meta["grid"], meta["with_res"]
        with_ln = ln_w is not None
        dy = dy.contiguous()
        M, C = x.shape
        ops.PROFILER.tag = ("K2", "bwd", M, C, (g.H // gs) * (g.W // gs))
        direct = _is_direct(meta, ctx.params) and meta["has_qkv_bias"]
        if direct:
            pl, pb, pwq, pbq, pwp, pbp = ctx.params
            dWq, dWp, dbq, dbp = pwq.grad.view(3 * C, C), pwp.grad.view(C, C), pbq.grad, pbp.grad
            dg, dbt = (pl.grad, pb.grad) if with_ln else (None, None)
        else:
            arena = _zeros(3 * C * C + C * C + 3 * C + C + 2 * C, x)
            off = 0
            dWq = arena[off:off + 3 * C * C].view(3 * C, C); off += 3 * C * C
            dWp = arena[off:off + C * C].view(C, C); off += C * C
            dbq = arena[off:off + 3 * C]; off += 3 * C
            dbp = arena[off:off + C]; off += C
            dg = arena[off:off + C]; off += C
            dbt = arena[off:off + C]; off += C
        gy = _scaled_grad(dy, scale, g.P, dbp)
        do = _empty((M, C), x)
        ops.gemm(gy, pp.wt, do)
        ops.wgrad(gy, o, dWp)
        dqkv = ops.grid_attn_bwd(qkv, do, g.B, g.H, g.W, C, heads, gs)
        dxn = _empty((M, C), x)
        if with_ln:
            ops.gemm(dqkv, pq.wt, dxn)
        else:
            ops.gemm(dqkv, pq.wt, dxn, residual=dy if with_res else None)
        ops.wgrad(dqkv, xn, dWq, bias_grad=dbq)
        if direct:
            dx = ops.layernorm_bwd(dxn, x, ln_w, mean, rstd, dy if with_res else None, dg, dbt) if with_ln else dxn
            _notify(ctx.params)
            return (dx,) + (None,) * 8
        dbq_out = dbq if meta["has_qkv_bias"] else None
        if with_ln:
            dx = ops.layernorm_bwd(dxn, x, ln_w, mean, rstd, dy if with_res else None, dg, dbt)
            return dx, dg, dbt, dWq, dbq_out, dWp, dbp, None, None
        return dxn, None, None, dWq, dbq_out, dWp, dbp, None, None


def grid_branch(x, ln_w, ln_b, wqkv, bqkv, wp, bp, scale, *, pq, pp, geom: Geom, heads: int, grid: int, eps: float,
                with_res: bool, capture=None, direct: bool = False) -> Tensor:
    meta = dict(pq=pq, pp=pp, geom=geom, heads=heads, grid=grid, eps=eps, with_res=with_res, capture=capture, direct=direct,
                has_qkv_bias=bqkv is not None)
    return GridBranchFn.apply(x, ln_w, ln_b, wqkv, bqkv, wp, bp, scale, meta)


# =================================================================================================
# MBConv (expand 1x1 + BN + act -> depthwise 3x3 + BN + act -> SE -> project 1x1 + BN, + residual)
# =================================================================================================
_CONV_STATS_FUSED = _os.environ.get("OGV_CONV_STATS_FUSED", "1") == "1"
_BN3_FUSED = _os.environ.get("OGV_BN3_STATS_FUSED", "1") == "1"  # 0: separate colstats pass (A/B: +0.12..0.18 ms per step)


class MBConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, we, g1, b1, wdw, g2, b2, sw1, sb1, sw2, sb2, wp, g3, b3, meta):
        pe: PreparedLinear = meta["pe"]
        ppj: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        act, training, eps, mom = meta["act"], meta["training"], meta["bn_eps"], meta["bn_momentum"]
        rm1, rv1, rm2, rv2, rm3, rv3 = meta["running"]
        x = x.contiguous()
        M, C = x.shape
        Cm = pe.w.shape[0]
        ops.PROFILER.tag = ("K3", "fwd", M, C)
        Cs = sw1.shape[0]
        # fp32 scratch for the three BatchNorms: [sum, sumsq, scale, shift, mean, rstd] x channels
        st = _scratch_zeros(6 * (2 * Cm + C), x)
        def carve(base, n):
            return [st[base + i * n: base + (i + 1) * n] for i in range(6)]
        s1 = carve(0, Cm)
        s2 = carve(6 * Cm, Cm)
        s3 = carve(12 * Cm, C)
        # expand
        e_pre = _empty((M, Cm), x)
        # BN1 batch statistics: a separate streaming pass.  Fusing them into this GEMM's epilogue was measured
        # SLOWER (358 us vs 192 + 107 us at stage 0): the epilogue warps of a C -> 4C GEMM are its critical path.
        # (The narrow project GEMM below and the convolutions of the conv -> BN -> act units do take their statistics
        # from the epilogue: 12 launches fewer per step, 0 .. 0.15 ms depending on the box.)
        ops.gemm(x, pe.w, e_pre)
        if training:
            ops.colstats(e_pre, s1[0], s1[1])
        ops.bn_finalize(s1[0], s1[1], g1, b1, rm1, rv1, s1[2], s1[3], s1[4], s1[5], M, eps, mom, training)
        # depthwise (BN1 + act on load, BN2 statistics on store)
        wdw2 = wdw.detach().reshape(Cm, 9).contiguous()
        d_pre = ops.dwconv_fwd(e_pre, s1[2], s1[3], wdw2, s2[0] if training else None, s2[1] if training else None,
                               g.B, g.H, g.W, act)
        ops.bn_finalize(s2[0], s2[1], g2, b2, rm2, rv2, s2[2], s2[3], s2[4], s2[5], M, eps, mom, training)
        # squeeze-excite on [B, Cm] rows: bf16 compute -> tcgen05 engine (the reference's autocast runs these 1x1
        # convs in bf16 too); fp32 compute -> fp32 FFMA engine.  The gate itself is always produced in fp32.
        pool = ops.se_pool(d_pre, s2[2], s2[3], g.B, g.P, act)
        ps1: PreparedLinear = meta["pse1"]
        ps2: PreparedLinear = meta["pse2"]
        cdt = x.dtype
        se_fused = ops.se_mlp_supported(Cm, Cs, cdt) and ps1.wt.is_contiguous() and ps2.wt.is_contiguous()
        if se_fused:  # both 1x1 convs, the casts and the two activations in one launch
            pool_c, s1_pre, s1a, gate_pre, gate = ops.se_mlp_fwd(pool, ps1.wt, sb1, ps2.wt, sb2, act)
        else:
            pool_c = pool if cdt == torch.float32 else ops.cast(pool, cdt)
            s1_pre = _empty((g.B, Cs), x, cdt)
            s1a = _empty((g.B, Cs), x, cdt)
            ops.gemm(pool_c, ps1.w, s1a, bias=sb1, pre_out=s1_pre, act=act)
            gate_pre = _empty((g.B, Cm), x, torch.float32)
            gate = _empty((g.B, Cm), x, torch.float32)
            ops.gemm(s1a, ps2.w, gate, bias=sb2, pre_out=gate_pre, act="sigmoid")
        d_act = ops.bn_act_gate(d_pre, s2[2], s2[3], gate, g.B, g.P, act)
        # project
        o_pre = _empty((M, C), x)
        if training and _BN3_FUSED and x.dtype == torch.bfloat16:
            # BN3 batch statistics from the narrow project GEMM's epilogue (its A operand, 4x wider than its output, is what
            # bounds it: the epilogue warps have the slack the C -> 4C expand GEMM's do not)
            ops.gemm(d_act, ppj.w, o_pre, col_sum=s3[0], col_sumsq=s3[1])
        else:
            ops.gemm(d_act, ppj.w, o_pre)
            if training:
                ops.colstats(o_pre, s3[0], s3[1])
        ops.bn_finalize(s3[0], s3[1], g3, b3, rm3, rv3, s3[2], s3[3], s3[4], s3[5], M, eps, mom, training)
        y = ops.bn_apply(o_pre, s3[2], s3[3], x if meta["use_res"] else None)
        ctx.meta = meta
        ctx.wdw2 = wdw2
        ctx.params = (we, g1, b1, wdw, g2, b2, sw1, sb1, sw2, sb2, wp, g3, b3)
        # `st` may be a slice of the engine's flat gradient/scratch arena, whose autograd version counter is bumped by
        # every in-place accumulation into ANY view of it: keep it out of save_for_backward's version check
        ctx.st = st
        ctx.save_for_backward(x, e_pre, d_pre, d_act, o_pre, pool_c, s1_pre, s1a, gate_pre, gate, g1, g2, g3, sw1, sw2)
        return y

    @staticmethod
    def backward(ctx, dy):
        meta = ctx.meta
        (x, e_pre, d_pre, d_act, o_pre, pool, s1_pre, s1a, gate_pre, gate, g1, g2, g3, sw1, sw2) = ctx.saved_tensors
        st = ctx.st
        pe: PreparedLinear = meta["pe"]
        ppj: PreparedLinear = meta["pp"]
        g: Geom = meta["geom"]
        act = meta["act"]
        dy = dy.contiguous()
        M, C = x.shape
        Cm = pe.w.shape[0]
        ops.PROFILER.tag = ("K3", "bwd", M, C)
        Cs = sw1.shape[0]
        def carve(base, n):
            return [st[base + i * n: base + (i + 1) * n] for i in range(6)]
        s1 = carve(0, Cm)
        s2 = carve(6 * Cm, Cm)
        s3 = carve(12 * Cm, C)
        sizes = [Cm * C, Cm, Cm, Cm * 9, Cm, Cm, Cs * Cm, Cs, Cm * Cs, Cm, C * Cm, C, C]
        direct = _is_direct(meta, ctx.params)
        if direct:  # every parameter starts on an 8-float granule of the flat gradient arena
            sl = [p.grad.view(-1) for p in ctx.params]
        else:
            # pad every slice to a multiple of 8 floats so vector loads of the BN partial sums stay aligned
            offs, tot = [], 0
            for s in sizes:
                offs.append(tot)
                tot += (s + 7) // 8 * 8
            arena = _zeros(tot, x)
            sl = [arena[o:o + s] for o, s in zip(offs, sizes)]
        dWe, dg1, db1, dwdw, dg2, db2, dsw1, dsb1, dsw2, dsb2, dWp, dg3, db3 = sl
        dWe = dWe.view(Cm, C); dsw1 = dsw1.view(Cs, Cm); dsw2 = dsw2.view(Cm, Cs); dWp = dWp.view(C, Cm)
        # BN3 backward
        # eval mode (running statistics): mean / variance are constants, so the batch-coupling terms of the
        # BatchNorm backward vanish -- same kernels, zero vectors in place of (dgamma, dbeta) in the apply steps;
        # the parameter gradients dgamma / dbeta themselves are unchanged.
        training = meta["training"]
        zC = None if training else _scratch_zeros(C, x)
        zM = None if training else _scratch_zeros(Cm, x)
        ops.bn_bwd_reduce(dy, o_pre, s3[4], s3[5], dg3, db3)
        do_pre = ops.bn_bwd_apply(dy, o_pre, s3[4], s3[5], g3, dg3 if training else zC, db3 if training else zC)
        # project backward
        dd_act = _empty((M, Cm), x)
        ops.gemm(do_pre, ppj.wt, dd_act)
        ops.wgrad(do_pre, d_act, dWp)
        # squeeze-excite + BN2 backward: one pass over (dd_act, d_pre) yields dgate and the per-image pieces of
        # the BN2 reductions; the tiny SE products then give dpool, and a [B, Cm] kernel finishes dgamma2 / dbeta2
        stats = ops.mbconv_bwd_stats(dd_act, d_pre, s2[2], s2[3], s2[4], s2[5], g.B, g.P, act)
        ps1: PreparedLinear = meta["pse1"]
        ps2: PreparedLinear = meta["pse2"]
        cdt = x.dtype
        if ops.se_mlp_supported(Cm, Cs, cdt) and ps1.w.is_contiguous() and ps2.w.is_contiguous():
            # sigma', both dgrads and act' in one launch; the two bias gradients ride with the weight-gradient GEMMs
            dgate_c, ds1_pre, dpool = ops.se_mlp_bwd(stats[0], gate_pre, s1_pre, ps2.w, ps1.w, act)
            ops.wgrad(dgate_c, s1a, dsw2, bias_grad=dsb2)
            ops.wgrad(ds1_pre, pool, dsw1, bias_grad=dsb1)
        else:
            dgate_pre = ops.mul_dact(stats[0], gate_pre, "sigmoid")
            dgate_c = dgate_pre if cdt == torch.float32 else ops.cast(dgate_pre, cdt)
            ds1_pre = _empty((g.B, Cs), x, cdt)
            ops.gemm(dgate_c, ps2.wt, ds1_pre, dact_src=s1_pre, dact=act)
            ops.wgrad(dgate_c, s1a, dsw2)
            ops.colsum(dgate_pre, dsb2)
            dpool = _empty((g.B, Cm), x, torch.float32)
            ops.gemm(ds1_pre, ps1.wt, dpool)
            ops.wgrad(ds1_pre, pool, dsw1)
            ops.colsum(ds1_pre, dsb1)
        ops.mbconv_bn2_finalize(stats, gate, dpool, dg2, db2, g.B, g.P)
        dd_pre = ops.dw_bn2_bwd_apply(dd_act, d_pre, gate, dpool, s2[2], s2[3], s2[4], s2[5], g2,
                                      dg2 if training else zM, db2 if training else zM, g.B, g.P, act)
        # depthwise backward (+ activation derivative, + BN1 reductions)
        du1 = ops.dwconv_bwd(dd_pre, e_pre, s1[2], s1[3], s1[4], s1[5], ctx.wdw2, dwdw, dg1, db1, g.B, g.H, g.W, act)
        de_pre = ops.bn_bwd_apply(du1, e_pre, s1[4], s1[5], g1, dg1 if training else zM, db1 if training else zM)
        # expand backward (+ skip-connection gradient)
        dx = _empty((M, C), x)
        ops.gemm(de_pre, pe.wt, dx, residual=dy if meta["use_res"] else None)
        ops.wgrad(de_pre, x, dWe)
        if direct:
            _notify(ctx.params)
            return (dx,) + (None,) * 14
        return (dx, dWe.view(meta["we_shape"]), dg1, db1, dwdw.view(meta["wdw_shape"]), dg2, db2,
                dsw1.view(meta["sw1_shape"]), dsb1, dsw2.view(meta["sw2_shape"]), dsb2, dWp.view(meta["wp_shape"]),
                dg3, db3, None)


def mbconv(x, we, g1, b1, wdw, g2, b2, sw1, sb1, sw2, sb2, wp, g3, b3, *, pe, pp, pse1, pse2, geom: Geom, act: str,
           training: bool, running, bn_eps: float, bn_momentum: float, use_res: bool, direct: bool = False) -> Tensor:
    meta = dict(pe=pe, pp=pp, pse1=pse1, pse2=pse2, geom=geom, act=act, training=training, running=running, bn_eps=bn_eps,
                direct=direct,
                bn_momentum=bn_momentum, use_res=use_res, we_shape=tuple(we.shape), wdw_shape=tuple(wdw.shape),
                sw1_shape=tuple(sw1.shape), sw2_shape=tuple(sw2.shape), wp_shape=tuple(wp.shape))
    return MBConvFn.apply(x, we, g1, b1, wdw, g2, b2, sw1, sb1, sw2, sb2, wp, g3, b3, meta)


# =================================================================================================
# stand-alone SqueezeExcite (mbc_conv.py:9-27) and the outlook core as their own autograd nodes
# =================================================================================================
class SqueezeExciteFn(torch.autograd.Function):
    """y = x * sigmoid(W2 act(W1 mean_hw(x) + b1) + b2) on rows [B*HW, C]; backward as autograd derives it:
    dx = dy*gate + dpool/HW, dgate = sum_hw dy*x (the same kernels MBConv's fused backward uses, with the BatchNorm of
    the depthwise stage replaced by the identity)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, B, HW, act):
        x = x.contiguous()
        M, C = x.shape
        Cs = w1.shape[0]
        dev = x.device
        one, zero = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        pool = ops.se_pool(x, one, zero, B, HW, "none")
        w1m, w2m = w1.detach().reshape(Cs, C), w2.detach().reshape(C, Cs)
        s1_pre = torch.empty((B, Cs), device=dev)
        s1a = torch.empty((B, Cs), device=dev)
        ops.gemm(pool, w1m, s1a, bias=b1.detach(), pre_out=s1_pre, act=act, engine=ENGINE_SIMT)
        gate_pre = torch.empty((B, C), device=dev)
        gate = torch.empty((B, C), device=dev)
        ops.gemm(s1a, w2m, gate, bias=b2.detach(), pre_out=gate_pre, act="sigmoid", engine=ENGINE_SIMT)
        y = ops.bn_act_gate(x, one, zero, gate, B, HW, "none")
        ctx.geom = (B, HW, act)
        ctx.shapes = (tuple(w1.shape), tuple(w2.shape))
        ctx.save_for_backward(x, pool, s1_pre, s1a, gate_pre, gate, w1m, w2m, one, zero)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pool, s1_pre, s1a, gate_pre, gate, w1m, w2m, one, zero = ctx.saved_tensors
        B, HW, act = ctx.geom
        dy = dy.contiguous()
        C, Cs = w2m.shape
        dgate = ops.se_bwd_reduce(dy, x, one, zero, B, HW, "none")
        dgate_pre = ops.mul_dact(dgate, gate_pre, "sigmoid")
        ds1_pre = torch.empty((B, Cs), device=x.device)
        ops.gemm(dgate_pre, w2m.t(), ds1_pre, dact_src=s1_pre, dact=act, engine=ENGINE_SIMT)
        dW2 = torch.zeros((C, Cs), device=x.device)
        ops.wgrad(dgate_pre, s1a, dW2, engine=ENGINE_SIMT)
        db2 = torch.zeros(C, device=x.device)
        ops.colsum(dgate_pre, db2)
        dpool = torch.empty((B, C), device=x.device)
        ops.gemm(ds1_pre, w1m.t(), dpool, engine=ENGINE_SIMT)
        dW1 = torch.zeros((Cs, C), device=x.device)
        ops.wgrad(ds1_pre, pool, dW1, engine=ENGINE_SIMT)
        db1 = torch.zeros(Cs, device=x.device)
        ops.colsum(ds1_pre, db1)
        # dx = (dy * gate + dpool / HW) through an identity BatchNorm (gamma = rstd = 1, no batch coupling)
        dx = ops.dw_bn2_bwd_apply(dy, x, gate, dpool, one, zero, zero, one, one, zero, zero, B, HW, "none")
        return dx, dW1.view(ctx.shapes[0]), db1, dW2.view(ctx.shapes[1]), db2, None, None, None


def squeeze_excite(x: Tensor, w1, b1, w2, b2, B: int, HW: int, act: str) -> Tensor:
    return SqueezeExciteFn.apply(x, w1, b1, w2, b2, B, HW, act)


class OutlookCoreFn(torch.autograd.Function):
    """softmax over the 9 taps + weighted 3x3 gather on va = [v | logits | pad] rows (outlook_attention.py:104-120)."""

    @staticmethod
    def forward(ctx, va, B, H, W, C, heads):
        va = va.contiguous()
        ctx.geom = (B, H, W, C, heads)
        ctx.save_for_backward(va)
        return ops.outlook_core_fwd(va, B, H, W, C, heads)

    @staticmethod
    def backward(ctx, dy):
        (va,) = ctx.saved_tensors
        B, H, W, C, heads = ctx.geom
        return ops.outlook_core_bwd(va, dy.contiguous(), B, H, W, C, heads), None, None, None, None, None


# =================================================================================================
# stand-alone LayerNorm on rows (LayerNorm2d) and layout conversion
# =================================================================================================
# =================================================================================================
# conv -> BatchNorm -> act units around the blocks: ConvStem (stem_head.py:23-32) and Downsample (downsampling.py:28-65)
# =================================================================================================
class ConvBnActFn(torch.autograd.Function):
    """conv -> BatchNorm -> act.  Both convolutions of the callers are patches x tcgen05 GEMM, forward, weight gradient
    and input gradient: the stem (3x3, stride 1, a few input channels: ogv_im2col3x3, one K tile, no input gradient) and
    the Downsample convolutions (downsampling.py:41-47: 3x3 stride 2, C -> 2C: ogv_im2col3x3_vec; the input gradient is
    dcols = dpre x W2 on the GEMM followed by the ogv_col2im3x3_vec gather); the 1x1 convolution of the "pool" kind is a
    plain GEMM on the rows.  Other geometries (kernel sizes, paddings, channel counts that are not a multiple of 8) go to
    the library convolution (OGV_CONV_LIB=1 forces it, for A/B measurements).  Everything after the convolution -- batch
    statistics, running-statistics update, normalise + activation, and the whole BatchNorm + activation backward -- runs
    on this package's streaming kernels over the channels_last rows, saving only the pre-BN conv output (the
    activation derivative is recomputed in both backward passes)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, meta):
        dt, act, training = meta["dtype"], meta["act"], meta["training"]
        stride, padding = meta["stride"], meta["padding"]
        rm, rv = meta["running"]
        xc = x.to(dt).contiguous(memory_format=torch.channels_last)
        Co, Cin, kh, kw = w.shape
        B = xc.shape[0]
        # patches x GEMM: "stem" when the patch row fits one K tile and nobody needs the input gradient (the network
        # input), "vec" for channel counts that are a multiple of 8 (the Downsample convolutions), "1x1" = plain GEMM
        Hi, Wi = xc.shape[2], xc.shape[3]
        k3 = kh == 3 and kw == 3 and padding == [1, 1] and stride[0] == stride[1]
        own = _os.environ.get("OGV_CONV_LIB", "0") != "1"
        if k3 and stride == [1, 1] and 9 * Cin <= 64 and not x.requires_grad:
            route = "stem"
        elif own and k3 and stride[0] in (1, 2) and Cin % 8 == 0:
            route = "vec"
        elif own and kh == 1 and kw == 1 and stride == [1, 1] and padding == [0, 0] and Cin % 8 == 0:
            route = "1x1"
        else:
            route = "lib"
        ops.PROFILER.tag = ("F2", "fwd", B * Hi * Wi // (stride[0] * stride[1]), Co, kh * kw * Cin, B * Hi * Wi * Cin)
        if route != "lib":
            if route == "stem":
                H, W = Hi, Wi
                kpad = (9 * Cin + 7) // 8 * 8
                cols = ops.im2col3x3(xc, kpad)
                w2 = torch.zeros((Co, kpad), device=w.device, dtype=dt)
                w2[:, :9 * Cin] = w.detach().permute(0, 2, 3, 1).reshape(Co, 9 * Cin)
            elif route == "vec":
                H, W = (Hi - 1) // stride[0] + 1, (Wi - 1) // stride[0] + 1
                w2 = w.detach().permute(0, 2, 3, 1).reshape(Co, 9 * Cin).to(dt)
                if ops.conv3x3_supported(xc, Co, stride[0]):
                    route = "implicit"  # no patch matrix at all: 5-D TMA boxes of x feed the GEMM
                    cols = xc
                else:
                    cols = ops.im2col3x3_vec(xc, stride[0])
            else:
                H, W = Hi, Wi
                cols = xc.permute(0, 2, 3, 1).reshape(B * H * W, Cin)
                w2 = w.detach().reshape(Co, Cin).to(dt)
            M = B * H * W
            # batch statistics from the GEMM's epilogue (bf16 staged epilogue; the separate pass otherwise)
            fused_stats = training and _CONV_STATS_FUSED and dt == torch.bfloat16 and Co % 8 == 0
            st = _scratch_zeros(6 * Co, xc)
            if route == "implicit":
                rows = ops.conv3x3_fwd(xc, w2, stride[0], col_sum=st[:Co] if fused_stats else None,
                                       col_sumsq=st[Co:2 * Co] if fused_stats else None)
            else:
                rows = _empty((M, Co), cols)
                if fused_stats:
                    ops.gemm(cols, w2, rows, col_sum=st[:Co], col_sumsq=st[Co:2 * Co])
                else:
                    ops.gemm(cols, w2, rows)
            saved = (cols, w2)
        else:
            wc = w.detach().to(dt).contiguous(memory_format=torch.channels_last)
            y_pre = torch.ops.aten.convolution(xc, wc, None, stride, padding, [1, 1], False, [0, 0], 1)
            y_pre = y_pre.contiguous(memory_format=torch.channels_last)
            _, _, H, W = y_pre.shape
            M = B * H * W
            rows = y_pre.permute(0, 2, 3, 1).reshape(M, Co)
            saved = (xc, wc)
            fused_stats = False
            st = _scratch_zeros(6 * Co, rows)
        ssum, ssq, scale, shift, mean, rstd = (st[i * Co:(i + 1) * Co] for i in range(6))
        if training and not fused_stats:
            ops.colstats(rows, ssum, ssq)
        ops.bn_finalize(ssum, ssq, gamma, beta, rm, rv, scale, shift, mean, rstd, M, meta["eps"], meta["momentum"],
                        training)
        out = ops.bn_act_apply(rows, scale, shift, act)
        ops.PROFILER.tag = None
        ctx.meta = meta
        ctx.geom = (B, Co, H, W)
        ctx.in_hw = (Hi, Wi)
        ctx.route = route
        ctx.w_shape = (Co, Cin, kh, kw)
        ctx.x_needs_grad = x.requires_grad
        ctx.x_dtype = x.dtype
        ctx.st = st  # a view of the per-step scratch arena (shared version counter): kept as an attribute
        ctx.params = (w, gamma, beta)
        ctx.save_for_backward(saved[0], saved[1], rows, gamma)
        return out.view(B, H, W, Co).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        a, wc, rows, gamma = ctx.saved_tensors
        meta = ctx.meta
        B, Co, H, W = ctx.geom
        M = B * H * W
        st = ctx.st
        scale, shift, mean, rstd = (st[i * Co:(i + 1) * Co] for i in range(2, 6))
        dyr = dy.to(rows.dtype).contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(M, Co)
        ops.PROFILER.tag = ("F2", "bwd", M, Co, ctx.w_shape[1] * ctx.w_shape[2] * ctx.w_shape[3],
                            B * ctx.in_hw[0] * ctx.in_hw[1] * ctx.w_shape[1])
        # engine "direct" mode: the BatchNorm reductions and the weight gradient accumulate straight into the parameters'
        # views of the flat gradient arena (zeroed once per step), nothing is handed back to autograd
        direct = _is_direct(meta, ctx.params)
        pw, pg, pb = ctx.params
        if direct:
            dgamma, dbeta = pg.grad, pb.grad
        else:
            red = _scratch_zeros(2 * Co, rows)
            dgamma, dbeta = red[:Co], red[Co:]
        ops.bn_act_bwd_reduce(dyr, rows, scale, shift, mean, rstd, dgamma, dbeta, meta["act"])
        if meta["training"]:
            dpre = ops.bn_act_bwd_apply(dyr, rows, scale, shift, mean, rstd, gamma.detach(), dgamma, dbeta, meta["act"])
        else:
            # running statistics: BatchNorm is a fixed affine map, dpre = gamma*rstd * g (no batch-mean terms)
            zero = _zeros(Co, gamma)
            dpre = ops.bn_act_bwd_apply(dyr, rows, scale, shift, mean, rstd, gamma.detach(), zero, zero, meta["act"])
        if ctx.route != "lib":
            _, Cin, kh, kw = ctx.w_shape
            # the weight gradient in the GEMM's [Co, (ky, kx, c)] layout: that IS the storage order of a channels_last
            # weight, so in direct mode the GEMM accumulates into the arena view itself
            dw2 = None
            if direct and wc.shape[1] == kh * kw * Cin:
                wg = pw.grad.permute(0, 2, 3, 1)
                if wg.is_contiguous():
                    dw2 = wg.reshape(Co, kh * kw * Cin)
            dw_in_place = dw2 is not None
            if dw2 is None:
                dw2 = _zeros((Co, wc.shape[1]), gamma)
            # on the main stream: autograd accumulates the returned view right away
            if ctx.route == "implicit":
                ops.conv3x3_wgrad(a, dpre, dw2, meta["stride"][0])
            else:
                ops._wgrad(dpre, a, dw2)
            dw = dw2[:, :kh * kw * Cin].view(Co, kh, kw, Cin).permute(0, 3, 1, 2)
            dx = None
            if ctx.x_needs_grad:  # never on the "stem" route
                Hi, Wi = ctx.in_hw
                w2t = wc.t().contiguous()  # [K, Co]: dcols = dpre x W2
                dcols = _empty((M, wc.shape[1]), dpre)
                ops.gemm(dpre, w2t, dcols)
                dxr = dcols if ctx.route == "1x1" else ops.col2im3x3_vec(dcols, B, Hi, Wi, Cin, meta["stride"][0])
                dx = dxr.view(B, Hi, Wi, Cin).permute(0, 3, 1, 2)
                if dx.dtype != ctx.x_dtype:
                    dx = dx.to(ctx.x_dtype)
        else:
            dpre4 = dpre.view(B, H, W, Co).permute(0, 3, 1, 2)
            dx, dw, _ = torch.ops.aten.convolution_backward(dpre4, a, wc, None, meta["stride"], meta["padding"], [1, 1],
                                                            False, [0, 0], 1, [ctx.x_needs_grad, True, False])
            dw = dw.float()
            if dx is not None and dx.dtype != ctx.x_dtype:
                dx = dx.to(ctx.x_dtype)
        ops.PROFILER.tag = None
        if direct:
            if not (ctx.route != "lib" and dw_in_place):
                pw.grad.add_(dw.to(pw.grad.dtype))
            _notify(ctx.params)
            return dx, None, None, None, None
        return dx, dw, dgamma.clone(), dbeta.clone(), None


def conv_bn_act(x: Tensor, w: Tensor, gamma: Tensor, beta: Tensor, *, running, stride: int, padding: int, act: str,
                eps: float, momentum: float, training: bool, dtype: torch.dtype, direct: bool = False) -> Tensor:
    meta = dict(running=running, stride=[stride, stride], padding=[padding, padding], act=act, eps=eps,
                momentum=momentum, training=training, dtype=dtype, direct=direct)
    return ConvBnActFn.apply(x, w, gamma, beta, meta)



# =================================================================================================
# classifier head: BatchNorm2d -> global average pool -> Linear  (Model_A_OutGridNet.py:52-53,65-67)
# =================================================================================================
class HeadFn(torch.autograd.Function):
    """logits = Linear(mean_hw(BatchNorm(x))) on rows [B*HW, C].  Statistics (ogv_colstats / ogv_bn_finalize), the
    normalise + pool as ONE pass (ogv_se_pool with the identity activation: the normalised map is never written), the
    classifier and its three gradients on the GEMM engine, BatchNorm backward on the streaming kernels."""

    @staticmethod
    def forward(ctx, rows, gamma, beta, wc, bc, meta):
        B, HW, training, dt = meta["B"], meta["HW"], meta["training"], meta["dtype"]
        rm, rv = meta["running"]
        rows = rows.contiguous()
        M, C = rows.shape
        K = wc.shape[0]
        ops.PROFILER.tag = ("HEAD", "fwd", M, C)
        st = _scratch_zeros(6 * C, rows)
        ssum, ssq, scale, shift, mean, rstd = (st[i * C:(i + 1) * C] for i in range(6))
        if training:
            ops.colstats(rows, ssum, ssq)
        ops.bn_finalize(ssum, ssq, gamma, beta, rm, rv, scale, shift, mean, rstd, M, meta["eps"], meta["momentum"], training)
        pool = ops.se_pool(rows, scale, shift, B, HW, "none")          # [B, C] fp32
        pool_c = pool if dt == torch.float32 else pool.to(dt)
        w_c = wc.detach() if dt == torch.float32 else wc.detach().to(dt)
        logits = _empty((B, K), rows, dt)
        ops.gemm(pool_c, w_c, logits, bias=bc.detach() if bc is not None else None)
        ops.PROFILER.tag = None
        ctx.meta = meta
        ctx.st = st
        ctx.has_bias = bc is not None
        ctx.params = (gamma, beta, wc, bc)
        ctx.save_for_backward(rows, gamma, pool_c, w_c)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        rows, gamma, pool_c, w_c = ctx.saved_tensors
        meta = ctx.meta
        B, HW, training = meta["B"], meta["HW"], meta["training"]
        M, C = rows.shape
        K = w_c.shape[0]
        st = ctx.st
        mean, rstd = st[4 * C:5 * C], st[5 * C:6 * C]
        ops.PROFILER.tag = ("HEAD", "bwd", M, C)
        dl = dlogits.to(pool_c.dtype).contiguous()
        pg, pb, pwc, pbc = ctx.params
        direct = _is_direct(meta, ctx.params) and pwc.grad.is_contiguous()
        dw = pwc.grad if direct else _zeros((K, C), gamma)
        ops._wgrad(dl, pool_c, dw)
        db = dlogits.float().sum(0) if ctx.has_bias else None  # [B, classes]: a few hundred KB
        dpool = _empty((B, C), rows, torch.float32)
        ops.gemm(dl, w_c.t().contiguous(), dpool)
        # every position of an image receives dpool / HW
        dyr = (dpool * (1.0 / HW)).to(rows.dtype).repeat_interleave(HW, dim=0)
        if direct:
            dgamma, dbeta = pg.grad, pb.grad
        else:
            red = _scratch_zeros(2 * C, rows)
            dgamma, dbeta = red[:C], red[C:]
        ops.bn_bwd_reduce(dyr, rows, mean, rstd, dgamma, dbeta)
        zero = None if training else _zeros(C, gamma)
        dx = ops.bn_bwd_apply(dyr, rows, mean, rstd, gamma.detach(), dgamma if training else zero, dbeta if training else zero)
        ops.PROFILER.tag = None
        if direct:
            if db is not None:
                pbc.grad.add_(db)
            _notify(ctx.params)
            return dx, None, None, None, None, None
        return dx, dgamma.clone(), dbeta.clone(), dw, db, None


def head(rows: Tensor, gamma: Tensor, beta: Tensor, wc: Tensor, bc: Optional[Tensor], *, B: int, HW: int, running,
         eps: float, momentum: float, training: bool, dtype: torch.dtype, direct: bool = False) -> Tensor:
    meta = dict(B=B, HW=HW, running=running, eps=eps, momentum=momentum, training=training, dtype=dtype, direct=direct)
    return HeadFn.apply(rows, gamma, beta, wc, bc, meta)


# =================================================================================================
# training criterion: nn.CrossEntropyLoss(label_smoothing=eps), mean reduction (train_full_model.py:52)
# =================================================================================================
class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, smoothing):
        logits = logits if logits.stride(-1) == 1 else logits.contiguous()
        loss = _zeros(1, logits)
        lse = ops.xent_fwd(logits, labels, smoothing, loss)
        ctx.smoothing = smoothing
        ctx.save_for_backward(logits, labels, lse)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        logits, labels, lse = ctx.saved_tensors
        g = gout.detach().float().reshape(1).contiguous()
        return ops.xent_bwd(logits, labels, lse, g, ctx.smoothing), None, None


def cross_entropy(logits: Tensor, target: Tensor, label_smoothing: float = 0.0) -> Tensor:
    """`F.cross_entropy(logits, target, label_smoothing=...)` with mean reduction on [B, K] logits and int64 class
    indices, as one kernel per direction (ogv_xent_fwd / ogv_xent_bwd).  Anything else the PyTorch function accepts
    (class weights, ignore_index, probabilities as targets, other reductions, CPU tensors) goes to PyTorch."""
    if (logits.is_cuda and logits.dim() == 2 and target.dim() == 1 and target.dtype == torch.int64
            and logits.dtype in (torch.float32, torch.bfloat16)):
        return CrossEntropyFn.apply(logits, target.contiguous(), float(label_smoothing))
    return torch.nn.functional.cross_entropy(logits, target, label_smoothing=label_smoothing)


class LayerNormRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, w, b, eps)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        C = x.shape[1]
        acc = _zeros(2 * C, x)
        dx = ops.layernorm_bwd(dy.contiguous(), x, w, mean, rstd, None, acc[:C], acc[C:])
        return dx, acc[:C], acc[C:], None


def layernorm_rows(x, w, b, eps: float) -> Tensor:
    return LayerNormRowsFn.apply(x, w, b, eps)


class NchwToRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return ops.nchw_to_rows(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        B, C, H, W = ctx.shape
        return ops.rows_to_nchw(dy.contiguous(), B, C, H, W)


def to_rows(x: Tensor) -> Tuple[Tensor, Geom]:
    """Logical NCHW tensor -> ([B*H*W, C] rows, geometry).  channels_last inputs are a free view."""
    if x.dim() != 4:
        raise ValueError(f"Expected a [B,C,H,W] tensor, got shape {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("outlook_grid_vision_transformer_b200 modules run on CUDA (sm_100a) only; "
                           f"got a tensor on '{x.device}'. There is no CPU fallback.")
    B, C, H, W = x.shape
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous():
        return xp.reshape(B * H * W, C), Geom(B, H, W)
    return NchwToRowsFn.apply(x), Geom(B, H, W)


def from_rows(y: Tensor, geom: Geom) -> Tensor:
    """rows -> logical NCHW view (channels_last strides; no copy)."""
    return y.view(geom.B, geom.H, geom.W, y.shape[1]).permute(0, 3, 1, 2)
