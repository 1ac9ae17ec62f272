"""nn.Module surface of the OutGridBlock hot path, mirroring the reference constructors, forward
signatures, attribute names and state_dict keys, with every forward routed
through the libogvit CUDA kernels.  Parameters live in ordinary nn.Conv2d / nn.Linear /
nn.LayerNorm / nn.BatchNorm2d containers so that default initialisation, `state_dict()` and the
weight-decay grouping of the reference training loop are unchanged; those containers' own
`forward` is never used on the fused path.

Reference files mirrored: src/model/outlook_attention.py, Outlook_Block.py, mbc_conv.py,
grid_partition.py, grid_attention.py, Out_Grid_Block.py, Grid_Only_Block.py, src/stage_config.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import torch
import torch.nn as nn

from . import functional as OF
from . import ops
from .functional import Geom

Tensor = torch.Tensor

_ACT_NAMES = {nn.SiLU: "silu", nn.ReLU: "relu", nn.GELU: "gelu"}


def make_activation(act: str) -> nn.Module:
    """silu | relu | gelu (exact erf).  (outlook_attention.py:6-14)"""
    name = act.lower()
    if name == "silu":
        return nn.SiLU(inplace=True)
    if name == "relu":
        return nn.ReLU(inplace=True)
    if name == "gelu":
        return nn.GELU()
    raise ValueError(f"Unknown activation '{act}'. Use one of: silu|gelu|relu")


def _act_name(m: nn.Module) -> str:
    for cls, name in _ACT_NAMES.items():
        if isinstance(m, cls):
            return name
    raise ValueError(f"unsupported activation module {type(m).__name__}")


def _compute_dtype(x: Tensor) -> torch.dtype:
    """bf16 under torch.autocast(cuda, bf16) or for bf16 inputs; fp32 otherwise."""
    if torch.is_autocast_enabled("cuda") if _NEW_AUTOCAST_API else torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda") if _NEW_AUTOCAST_API else torch.get_autocast_gpu_dtype()
        if dt == torch.bfloat16:
            return torch.bfloat16
        raise NotImplementedError(f"autocast dtype {dt} is not supported on the B200 path; use bfloat16")
    if x.dtype in (torch.float32, torch.bfloat16):
        return x.dtype
    raise TypeError(f"unsupported input dtype {x.dtype}; use float32 or bfloat16")


try:
    torch.is_autocast_enabled("cuda")
    _NEW_AUTOCAST_API = True
except TypeError:  # pragma: no cover - older torch
    _NEW_AUTOCAST_API = False


FORCE_PREP = False  # set while a CUDA graph is being captured: the casts must be graph nodes
# Bumped by every optimizer update the package performs (engine.TrainStep, eager or graph replay).  Part of the
# eval-mode cache key below: a fused optimizer (ours, and torch's fused AdamW) updates the fp32 master weights
# without touching tensor._version, so "same storage, same version" does NOT mean "same values".
PARAM_GENERATION = 0


def note_parameters_updated() -> None:
    """Call after updating parameters in a way autograd's version counter cannot see (fused / graph-replayed
    optimizers): invalidates every cached compute-dtype weight copy used by eval-mode forwards."""
    global PARAM_GENERATION
    PARAM_GENERATION += 1


class _Prep:
    """Compute-dtype weight copies (what autocast's weight cast produces).  In training mode they are
    re-derived on EVERY forward: an optimizer may update the fp32 master weights without touching the
    tensor version counter (torch's fused AdamW does), so no cache key is trustworthy there.  Two ways:
    per layer (a cast launch per weight, the default), or in bulk -- the casts of the last per-layer build are
    remembered (`_bulk`), refresh_prepared(model) re-runs all of them in ONE launch at the start of a step and,
    while `_BULK_FRESH` is set, get() hands out those refreshed buffers.  In eval mode the copies are cached,
    keyed by storage / version / dtype."""

    def __init__(self):
        self._key = None
        self._val = None
        self._bulk = None  # (key, value, [(src, dst, dst_t), ...]) of the last training-mode build

    @staticmethod
    def _bulk_key(params, dtype):
        return (dtype,) + tuple(p.data_ptr() if p is not None else None for p in params)

    def get(self, training, params, dtype, build):
        if training or FORCE_PREP:
            self._key = None  # the optimizer step that follows makes this copy stale: never reuse it
            if _BULK_FRESH and self._bulk is not None and self._bulk[0] == self._bulk_key(params, dtype):
                return self._bulk[1]
            OF._CAST_LOG = log = []
            try:
                val = build()
            finally:
                OF._CAST_LOG = None
            self._bulk = (self._bulk_key(params, dtype), val, log) if log else None
            return val
        key = (dtype, PARAM_GENERATION) + tuple((p.data_ptr(), p._version) if p is not None else None for p in params)
        if key != self._key:
            self._val = build()
            self._key = key
        return self._val


_BULK_FRESH = False  # set by the train-step executor between refresh_prepared() and the end of the forward pass


def refresh_prepared(model: nn.Module) -> bool:
    """Re-derive every compute-dtype weight copy recorded by the model's training-mode forward in ONE launch
    (ogv_cast_batch) instead of one launch per weight.  Returns False when there is nothing recorded yet (first
    step): the caller then leaves `_BULK_FRESH` unset and the layers cast per weight as usual."""
    preps = [p for m in model.modules() for p in m.__dict__.get("_ogv_prep", {}).values() if p._bulk is not None]
    if not preps:
        return False
    sig = tuple(id(p._bulk) for p in preps)
    cached = model.__dict__.get("_ogv_cast_batch")
    if cached is None or cached[0] != sig:
        triples = [t for p in preps for t in p._bulk[2]]
        cached = (sig, ops.CastBatch(triples), [p._bulk for p in preps])  # keeps the ids in `sig` alive
        model.__dict__["_ogv_cast_batch"] = cached
    cached[1].run()
    return True


def _prep_attr(mod: nn.Module, name: str) -> _Prep:
    store = mod.__dict__.setdefault("_ogv_prep", {})
    if name not in store:
        store[name] = _Prep()
    return store[name]


# Set by the train-step executor: an iterator over the rows of ONE pre-drawn [n_droppath, B] scale table (a single
# Bernoulli launch per step instead of four launches per DropPath); rows are consumed in call order.
DROP_TABLE = None


def _direct(mod: nn.Module) -> bool:
    """Parameter gradients of this module go straight into the engine's flat gradient arena (functional._is_direct)."""
    return bool(mod.__dict__.get("_ogv_direct", False)) and torch.is_grad_enabled()


def _drop_scale(dp: nn.Module, x: Tensor, batch: int) -> Optional[Tensor]:
    """Per-sample stochastic-depth scale, drawn exactly like DropPath.forward (Outlook_Block.py:15-22)."""
    if not isinstance(dp, DropPath) or dp.drop_prob == 0.0 or not dp.training:
        return None
    if DROP_TABLE is not None:
        return next(DROP_TABLE)
    keep = 1.0 - dp.drop_prob
    mask = torch.empty((batch, 1, 1, 1), device=x.device, dtype=x.dtype).bernoulli_(keep)
    return (mask.to(torch.float32) / keep).reshape(batch).contiguous()


# =================================================================================================
# outlook_attention.py
# =================================================================================================
class LayerNorm2d(nn.Module):
    """LayerNorm over C at every (h, w) of a [B,C,H,W] tensor.  (outlook_attention.py:17-31)"""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True):
        super().__init__()
        self.ln = nn.LayerNorm(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        rows = rows.to(_compute_dtype(x)) if rows.dtype != _compute_dtype(x) else rows
        w, b = self.ln.weight, self.ln.bias
        if w is None:
            w = torch.ones(rows.shape[1], device=x.device)
            b = torch.zeros(rows.shape[1], device=x.device)
        return OF.from_rows(OF.layernorm_rows(rows, w, b, self.ln.eps), geom)


class MLP2d(nn.Module):
    """1x1 conv -> act -> 1x1 conv on NCHW.  (outlook_attention.py:33-49)"""

    def __init__(self, dim, mlp_ratio=4.0, drop=0.0, act="gelu"):
        super().__init__()
        hidden = max(1, int(dim * mlp_ratio))
        self.fc1 = nn.Conv2d(dim, hidden, 1)
        self.act = make_activation(act)
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Conv2d(hidden, dim, 1)
        self.drop2 = nn.Dropout(drop)

    def _prepared(self, dtype):
        p1 = _prep_attr(self, "fc1").get(self.training, [self.fc1.weight], dtype, lambda: OF.prepare_linear(self.fc1.weight, dtype))
        p2 = _prep_attr(self, "fc2").get(self.training, [self.fc2.weight], dtype, lambda: OF.prepare_linear(self.fc2.weight, dtype))
        return p1, p2

    def _check_dropout(self):
        if self.training and (self.drop1.p > 0 or self.drop2.p > 0):
            raise NotImplementedError("ffn dropout > 0 is not implemented on the fused path (all configs use 0.0)")

    def rows_forward(self, rows: Tensor, geom: Geom, ln: Optional[nn.LayerNorm], scale: Optional[Tensor],
                     with_res: bool) -> Tensor:
        self._check_dropout()
        p1, p2 = self._prepared(rows.dtype)
        return OF.mlp_branch(rows, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                             self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, scale, p1=p1, p2=p2,
                             eps=ln.eps if ln is not None else 0.0, act=_act_name(self.act),
                             rows_per_sample=geom.P, with_res=with_res, direct=_direct(self))

    def forward(self, x):
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return OF.from_rows(self.rows_forward(rows, geom, None, None, False), geom)


class OutlookAttention2d(nn.Module):
    """Outlook attention on NCHW: per-position k*k weights -> softmax -> weighted 3x3 gather of v -> proj.
    (outlook_attention.py:52-124; forward is a gather, no fold -- SURVEY fact 2)"""

    def __init__(self, dim: int, num_heads: int = 6, kernel_size: int = 3, stride: int = 1, attn_drop: float = 0.0,
                 proj_drop: float = 0.0, qkv_bias: bool = True):
        super().__init__()
        assert dim % num_heads == 0, "dim must be divisible by num_heads"
        if kernel_size <= 0 or kernel_size % 2 == 0:
            raise ValueError("kernel_size must be odd and >0 (e.g., 3,5,7)")
        if stride <= 0:
            raise ValueError("stride must be > 0")
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.kernel_size = kernel_size
        self.stride = stride
        kk = kernel_size * kernel_size
        bias = bool(qkv_bias)
        self.attn = nn.Conv2d(dim, num_heads * kk, kernel_size=1, bias=bias)
        self.v = nn.Conv2d(dim, dim, kernel_size=1, bias=bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1, bias=True)
        self.proj_drop = nn.Dropout(proj_drop)

    def _prepared(self, dtype):
        pva, bva = _prep_attr(self, "va").get(
            self.training, [self.v.weight, self.v.bias, self.attn.weight, self.attn.bias], dtype,
            lambda: OF.prepare_outlook_va(self.v.weight, self.v.bias, self.attn.weight, self.attn.bias, dtype))
        pp = _prep_attr(self, "proj").get(self.training, [self.proj.weight], dtype, lambda: OF.prepare_linear(self.proj.weight, dtype))
        return pva, bva, pp

    def _hooked(self) -> bool:
        """A forward (pre-)hook sits on one of the 1x1 convs -- e.g. the logits capture of the attention-map tools
        (src/experiments/heat_map_att_outlooker.py:25-42), which hooks `self.attn`."""
        return any(m._forward_hooks or m._forward_pre_hooks for m in (self.attn, self.v, self.proj))

    def _composed_rows_forward(self, rows, geom, ln, scale, with_res):
        """The same math with the three 1x1 convs CALLED as modules, so hooks on them fire with the tensors the
        reference would hand them (NCHW input, NCHW output); LayerNorm and the softmax / gather core stay on the
        library kernels.  Analysis path: correctness over speed."""
        C, nl = self.dim, self.attn.weight.shape[0]
        xn = OF.layernorm_rows(rows, ln.weight, ln.bias, ln.eps) if ln is not None else rows
        xn4 = OF.from_rows(xn, geom)
        with torch.autocast("cuda", enabled=False):
            w_dt = xn4.dtype
            conv = lambda m, t: m(t) if w_dt == torch.float32 else m(t.float()).to(w_dt)  # noqa: E731
            logits = conv(self.attn, xn4)                      # hook on self.attn sees [B, heads*9, H, W]
            v = conv(self.v, xn4)
            va = torch.zeros((rows.shape[0], OF.outlook_npad(C, self.num_heads)), device=rows.device, dtype=rows.dtype)
            va[:, :C] = v.permute(0, 2, 3, 1).reshape(-1, C)
            va[:, C:C + nl] = logits.permute(0, 2, 3, 1).reshape(-1, nl)
            yc = OF.OutlookCoreFn.apply(va, geom.B, geom.H, geom.W, C, self.num_heads)
            y = conv(self.proj, OF.from_rows(yc, geom)).permute(0, 2, 3, 1).reshape(-1, C)
        if scale is not None:
            y = y * scale.repeat_interleave(geom.P).to(y.dtype)[:, None]
        return rows + y if with_res else y

    def rows_forward(self, rows: Tensor, geom: Geom, ln: Optional[nn.LayerNorm], scale: Optional[Tensor],
                     with_res: bool) -> Tensor:
        if self.kernel_size == 3 and self.stride == 1 and self._hooked():
            return self._composed_rows_forward(rows, geom, ln, scale, with_res)
        if self.kernel_size != 3 or self.stride != 1:
            raise NotImplementedError("the sm_100a outlook kernel implements kernel_size=3, stride=1 "
                                      "(the only configuration OutGridBlock constructs, Out_Grid_Block.py:44-49)")
        if self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0):
            raise NotImplementedError("attention/projection dropout > 0 is not implemented (all configs use 0.0)")
        pva, bva, pp = self._prepared(rows.dtype)
        return OF.outlook_branch(rows, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                                 self.v.weight, self.v.bias, self.attn.weight, self.attn.bias, self.proj.weight,
                                 self.proj.bias, scale, pva=pva, bva=bva, pp=pp, geom=geom, heads=self.num_heads,
                                 eps=ln.eps if ln is not None else 0.0, with_res=with_res, direct=_direct(self),
                                 flat=self.__dict__.get("_ogv_flat"))

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return OF.from_rows(self.rows_forward(rows, geom, None, None, False), geom)


# =================================================================================================
# Outlook_Block.py
# =================================================================================================
class _RowScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, rows_per_scale):
        ctx.save_for_backward(scale)
        ctx.rps = rows_per_scale
        return ops.rowscale(x.contiguous(), scale, rows_per_scale)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        return ops.rowscale(dy.contiguous(), scale, ctx.rps), None, None


class DropPath(nn.Module):
    """Stochastic depth: per-sample Bernoulli(keep) mask / keep.  (Outlook_Block.py:7-22)"""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x: Tensor) -> Tensor:
        if self.drop_prob == 0.0 or (not self.training):
            return x
        scale = _drop_scale(self, x, x.shape[0])
        flat = x.contiguous().view(x.shape[0], -1)
        if flat.shape[1] % 8 != 0:
            raise NotImplementedError("DropPath kernel needs numel/batch to be a multiple of 8")
        return _RowScaleFn.apply(flat, scale, 1).view(x.shape)


class OutlookerBlock2d(nn.Module):
    """x + dp1(attn(norm1 x)); x + dp2(mlp(norm2 x)) on NCHW.  (Outlook_Block.py:26-64)"""

    def __init__(self, dim: int, num_heads: int, kernel_size: int = 3, stride: int = 1, mlp_ratio: float = 2.0,
                 attn_drop: float = 0.0, proj_drop: float = 0.0, drop_path: float = 0.0, mlp_drop: float = 0.0,
                 act: str = "gelu", norm_eps: float = 1e-6):
        super().__init__()
        self.norm1 = LayerNorm2d(dim, eps=norm_eps)
        self.attn = OutlookAttention2d(dim=dim, num_heads=num_heads, kernel_size=kernel_size, stride=stride,
                                       attn_drop=attn_drop, proj_drop=proj_drop)
        self.dp1 = DropPath(drop_path) if drop_path > 0 else nn.Identity()
        self.norm2 = LayerNorm2d(dim, eps=norm_eps)
        self.mlp = MLP2d(dim=dim, mlp_ratio=mlp_ratio, drop=mlp_drop, act=act)
        self.dp2 = DropPath(drop_path) if drop_path > 0 else nn.Identity()

    def rows_forward(self, rows: Tensor, geom: Geom) -> Tensor:
        s1 = _drop_scale(self.dp1, rows, geom.B)
        rows = self.attn.rows_forward(rows, geom, self.norm1.ln, s1, True)
        s2 = _drop_scale(self.dp2, rows, geom.B)
        return self.mlp.rows_forward(rows, geom, self.norm2.ln, s2, True)

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return OF.from_rows(self.rows_forward(rows, geom), geom)


# =================================================================================================
# mbc_conv.py
# =================================================================================================
class SqueezeExcite(nn.Module):
    """Global-average squeeze, two 1x1 convs, sigmoid gate.  (mbc_conv.py:9-27)"""

    def __init__(self, channels: int, se_ratio: float = 0.25, act: str = "silu"):
        super().__init__()
        if not (0.0 < se_ratio <= 1.0):
            raise ValueError("se_ratio must be in (0, 1].")
        hidden = max(1, int(channels * se_ratio))
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, hidden, kernel_size=1, bias=True)
        self.act = make_activation(act)
        self.fc2 = nn.Conv2d(hidden, channels, kernel_size=1, bias=True)
        self.gate = nn.Sigmoid()

    def forward(self, x: Tensor) -> Tensor:
        """Stand-alone use (inside MBConv the SE is part of the fused autograd function).  fp32 or bf16 rows; the two
        tiny 1x1 convs on the pooled [B, C] vector run in fp32."""
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        y = OF.squeeze_excite(rows, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, geom.B, geom.P,
                              _act_name(self.act))
        return OF.from_rows(y, geom)


ActType = Literal["silu", "gelu", "relu"]


@dataclass(frozen=True)
class MBConvConfig:
    expand_ratio: float = 4.0
    se_ratio: float = 0.25
    act: ActType = "silu"
    use_bn: bool = True
    drop_path: float = 0.0


class MBConv(nn.Module):
    """Expand 1x1 -> depthwise 3x3 -> SE -> project 1x1, BatchNorm after each conv, residual when
    stride == 1 and in_ch == out_ch.  (mbc_conv.py:44-98)"""

    def __init__(self, in_ch: int, out_ch: int, stride: int = 1, cfg: MBConvConfig = MBConvConfig()):
        super().__init__()
        if in_ch <= 0 or out_ch <= 0:
            raise ValueError("in_ch and out_ch must be > 0")
        if stride not in (1, 2):
            raise ValueError("stride must be 1 or 2")
        self.in_ch = in_ch
        self.out_ch = out_ch
        self.stride = stride
        self._cfg = cfg
        norm = (lambda c: nn.BatchNorm2d(c)) if cfg.use_bn else (lambda c: nn.Identity())
        act = make_activation(cfg.act)
        mid = max(1, int(round(in_ch * cfg.expand_ratio)))
        if mid != in_ch:
            self.expand = nn.Sequential(nn.Conv2d(in_ch, mid, kernel_size=1, bias=not cfg.use_bn), norm(mid), act)
        else:
            self.expand = nn.Identity()
        self.depthwise = nn.Sequential(
            nn.Conv2d(mid, mid, kernel_size=3, stride=stride, padding=1, groups=mid, bias=not cfg.use_bn),
            norm(mid), act)
        self.se = SqueezeExcite(mid, se_ratio=cfg.se_ratio, act=cfg.act) if cfg.se_ratio > 0 else nn.Identity()
        self.project = nn.Sequential(nn.Conv2d(mid, out_ch, kernel_size=1, bias=not cfg.use_bn), norm(out_ch))
        self.use_res = (stride == 1 and in_ch == out_ch)
        self.drop_path = DropPath(cfg.drop_path) if (cfg.drop_path and cfg.drop_path > 0) else nn.Identity()

    def _check_supported(self):
        cfg = self._cfg
        if (isinstance(self.expand, nn.Identity) or isinstance(self.se, nn.Identity) or not cfg.use_bn
                or self.stride != 1 or isinstance(self.drop_path, DropPath)):
            raise NotImplementedError(
                "the sm_100a MBConv path implements the configuration OutGridBlock constructs "
                "(expand_ratio != 1, SE on, BatchNorm on, stride 1, drop_path 0; Out_Grid_Block.py:58-66)")

    def _prepared(self, dtype):
        we, wp = self.expand[0].weight, self.project[0].weight
        pe = _prep_attr(self, "expand").get(self.training, [we], dtype, lambda: OF.prepare_linear(we, dtype))
        pp = _prep_attr(self, "project").get(self.training, [wp], dtype, lambda: OF.prepare_linear(wp, dtype))
        w1, w2 = self.se.fc1.weight, self.se.fc2.weight
        ps1 = _prep_attr(self, "se1").get(self.training, [w1], dtype, lambda: OF.prepare_linear(w1, dtype))
        ps2 = _prep_attr(self, "se2").get(self.training, [w2], dtype, lambda: OF.prepare_linear(w2, dtype))
        return pe, pp, ps1, ps2

    def rows_forward(self, rows: Tensor, geom: Geom) -> Tensor:
        self._check_supported()
        bn1, bn2, bn3 = self.expand[1], self.depthwise[1], self.project[1]
        pe, pp, ps1, ps2 = self._prepared(rows.dtype)
        training = self.training and bn1.track_running_stats is not None
        if self.training and not self.__dict__.get("_ogv_nbt_shared", False):
            # (under engine.FlatState the counters of every MBConv are views of one int64 arena bumped once per step)
            for bn in (bn1, bn2, bn3):
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
        mom = bn1.momentum if bn1.momentum is not None else 0.1
        running = (bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var, bn3.running_mean,
                   bn3.running_var)
        return OF.mbconv(rows, self.expand[0].weight, bn1.weight, bn1.bias, self.depthwise[0].weight, bn2.weight,
                         bn2.bias, self.se.fc1.weight, self.se.fc1.bias, self.se.fc2.weight, self.se.fc2.bias,
                         self.project[0].weight, bn3.weight, bn3.bias, pe=pe, pp=pp, pse1=ps1, pse2=ps2, geom=geom,
                         act=_act_name(self.depthwise[2]), training=training, running=running, bn_eps=bn1.eps,
                         bn_momentum=mom, use_res=self.use_res, direct=_direct(self))

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return OF.from_rows(self.rows_forward(rows, geom), geom)


# =================================================================================================
# grid_partition.py -- kept for API compatibility (analysis tools); the fused path never calls them
# =================================================================================================
def grid_partition(x: Tensor, grid_size: int):
    """[B,H,W,C] -> ([B*g*g, H/g, W/g, C], meta): group (h%g, w%g), token (h//g, w//g).  (grid_partition.py:3-17)"""
    if x.ndim != 4:
        raise ValueError(f"Expected x.ndim==4 (BHWC). Got shape {tuple(x.shape)}")
    B, H, W, C = x.shape
    g = grid_size
    if g <= 0:
        raise ValueError("grid_size must be > 0")
    if (H % g) != 0 or (W % g) != 0:
        raise ValueError(f"H and W must be divisible by grid_size. Got H={H}, W={W}, g={g}")
    hg, wg = H // g, W // g
    grids = x.reshape(B, hg, g, wg, g, C).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, hg, wg, C)
    return grids, (B, H, W, C, g)


def grid_unpartition(grids: Tensor, meta) -> Tensor:
    """Inverse of grid_partition.  (grid_partition.py:20-32)"""
    if grids.ndim != 4:
        raise ValueError(f"Expected grids.ndim==4. Got shape {tuple(grids.shape)}")
    B, H, W, C, g = meta
    hg, wg = H // g, W // g
    if grids.shape[0] != B * g * g:
        raise ValueError(f"grids.shape[0] must be B*g*g = {B*g*g}. Got {grids.shape[0]}")
    if grids.shape[1] != hg or grids.shape[2] != wg or grids.shape[3] != C:
        raise ValueError(f"grids shape mismatch. Expected (*,{hg},{wg},{C}) got {tuple(grids.shape)}")
    return grids.reshape(B, g, g, hg, wg, C).permute(0, 3, 1, 4, 2, 5).reshape(B, H, W, C)


# =================================================================================================
# grid_attention.py
# =================================================================================================
AttnMode = Literal["grid"]


@dataclass(frozen=True)
class AttentionConfig:
    dim: int
    num_heads: int
    qkv_bias: bool = True
    attn_drop: float = 0.0
    proj_drop: float = 0.0


@dataclass(frozen=True)
class GridAttention2DConfig:
    mode: AttnMode
    dim: int
    num_heads: int
    grid_size: int
    window_size: int = 1
    qkv_bias: bool = True
    attn_drop: float = 0.0
    proj_drop: float = 0.0


class MultiHeadSelfAttention(nn.Module):
    """MHSA over token groups [Bgrp, N, C].  (grid_attention.py:33-89)"""

    def __init__(self, cfg: AttentionConfig):
        super().__init__()
        if cfg.dim <= 0:
            raise ValueError("cfg.dim must be > 0")
        if cfg.num_heads <= 0:
            raise ValueError("cfg.num_heads must be > 0")
        if cfg.dim % cfg.num_heads != 0:
            raise ValueError(f"dim ({cfg.dim}) must be divisible by num_heads ({cfg.num_heads})")
        self.dim = cfg.dim
        self.num_heads = cfg.num_heads
        self.head_dim = cfg.dim // cfg.num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(cfg.dim, 3 * cfg.dim, bias=cfg.qkv_bias)
        self.attn_drop = nn.Dropout(cfg.attn_drop)
        self.proj = nn.Linear(cfg.dim, cfg.dim, bias=True)
        self.proj_drop = nn.Dropout(cfg.proj_drop)

    def _prepared(self, dtype):
        pq = _prep_attr(self, "qkv").get(self.training, [self.qkv.weight], dtype, lambda: OF.prepare_linear(self.qkv.weight, dtype))
        pp = _prep_attr(self, "proj").get(self.training, [self.proj.weight], dtype, lambda: OF.prepare_linear(self.proj.weight, dtype))
        return pq, pp

    def _capture(self):
        if not getattr(self, "capture_attn", False):
            return None

        def store(attn: Tensor):
            self.last_attn = attn.detach()
            self.last_attn_postdrop = attn.detach()  # attn_drop is 0 on every supported config

        return store

    def rows_forward(self, rows: Tensor, geom: Geom, grid: int, ln: Optional[nn.LayerNorm], scale: Optional[Tensor],
                     with_res: bool) -> Tensor:
        if self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0):
            raise NotImplementedError("attention/projection dropout > 0 is not implemented (all configs use 0.0)")
        pq, pp = self._prepared(rows.dtype)
        return OF.grid_branch(rows, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                              self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, scale, pq=pq, pp=pp,
                              geom=geom, heads=self.num_heads, grid=grid, eps=ln.eps if ln is not None else 0.0,
                              with_res=with_res, capture=self._capture(), direct=_direct(self))

    def forward(self, x: Tensor) -> Tensor:
        if x.ndim != 3:
            raise ValueError(f"Expected x.ndim==3 with shape [B, N, C]. Got {tuple(x.shape)}")
        Bg, N, C = x.shape
        if C != self.dim:
            raise ValueError(f"Expected last dim C={self.dim}. Got C={C}")
        if not x.is_cuda:
            raise RuntimeError("MultiHeadSelfAttention runs on CUDA (sm_100a) only; there is no CPU fallback")
        dt = _compute_dtype(x)
        rows = x.reshape(Bg * N, C)
        rows = rows.to(dt) if rows.dtype != dt else rows
        # already-partitioned tokens: every sample is one group (g = 1, H = N, W = 1)
        y = self.rows_forward(rows, Geom(Bg, N, 1), 1, None, None, False)
        return y.view(Bg, N, C)


class GridAttention2D(nn.Module):
    """Grid attention on BHWC.  (grid_attention.py:93-131)"""

    def __init__(self, cfg: GridAttention2DConfig):
        super().__init__()
        if cfg.mode != "grid":
            raise ValueError("This minimal version only supports mode='grid'")
        self.cfg = cfg
        self.mhsa = MultiHeadSelfAttention(AttentionConfig(dim=cfg.dim, num_heads=cfg.num_heads, qkv_bias=cfg.qkv_bias,
                                                           attn_drop=cfg.attn_drop, proj_drop=cfg.proj_drop))

    def _note_meta(self, B, H, W, C):
        g = self.cfg.grid_size
        if g <= 0:
            raise ValueError("grid_size must be > 0")
        if (H % g) != 0 or (W % g) != 0:
            raise ValueError(f"H and W must be divisible by grid_size. Got H={H}, W={W}, g={g}")
        self._last_meta = (B, H, W, C, g)
        self._last_grid_hw = (H // g, W // g)
        self._last_g = g
        return g

    def rows_forward(self, rows: Tensor, geom: Geom, ln, scale, with_res: bool) -> Tensor:
        g = self._note_meta(geom.B, geom.H, geom.W, rows.shape[1])
        return self.mhsa.rows_forward(rows, geom, g, ln, scale, with_res)

    def forward(self, x: Tensor) -> Tensor:
        if x.ndim != 4:
            raise ValueError(f"Expected x.ndim==4 (BHWC). Got {tuple(x.shape)}")
        B, H, W, C = x.shape
        if C != self.cfg.dim:
            raise ValueError(f"Expected C=={self.cfg.dim}. Got C={C}")
        if not x.is_cuda:
            raise RuntimeError("GridAttention2D runs on CUDA (sm_100a) only; there is no CPU fallback")
        dt = _compute_dtype(x)
        rows = x.reshape(B * H * W, C)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return self.rows_forward(rows, Geom(B, H, W), None, None, False).view(B, H, W, C)


# =================================================================================================
# Out_Grid_Block.py / Grid_Only_Block.py
# =================================================================================================
class MLP(nn.Module):
    """Linear -> act -> Linear over the last dim.  (Out_Grid_Block.py:10-32)"""

    def __init__(self, dim: int, mlp_ratio: float = 4.0, drop: float = 0.0, act: str = "gelu"):
        super().__init__()
        hidden = max(1, int(dim * mlp_ratio))
        self.fc1 = nn.Linear(dim, hidden)
        self.act = make_activation(act)
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(drop)

    _prepared = MLP2d._prepared
    _check_dropout = MLP2d._check_dropout
    rows_forward = MLP2d.rows_forward

    def forward(self, x: Tensor) -> Tensor:
        if x.shape[-1] != self.fc1.in_features:
            raise ValueError(f"MLP expected last dim={self.fc1.in_features}, got {x.shape[-1]}")
        if not x.is_cuda:
            raise RuntimeError("MLP runs on CUDA (sm_100a) only; there is no CPU fallback")
        dt = _compute_dtype(x)
        rows = x.reshape(-1, x.shape[-1])
        rows = rows.to(dt) if rows.dtype != dt else rows
        y = self.rows_forward(rows, Geom(rows.shape[0], 1, 1), None, None, False)
        return y.view(x.shape)


class _GridMlpTail(nn.Module):
    """Shared tail of OutGridBlock / GridOnlyBlock: LN -> grid attention -> +res, LN -> MLP -> +res."""

    def _build_tail(self, cfg):
        C = cfg.dim
        self.norm2 = nn.LayerNorm(C)
        self.grid_attn = GridAttention2D(GridAttention2DConfig(
            mode="grid", dim=C, num_heads=cfg.num_heads, window_size=getattr(cfg, "window_size", 1),
            grid_size=cfg.grid_size, qkv_bias=True, attn_drop=cfg.attn_drop, proj_drop=cfg.proj_drop))
        self.dp2 = DropPath(cfg.drop_path) if cfg.drop_path > 0 else nn.Identity()
        self.norm3 = nn.LayerNorm(C)
        self.mlp = MLP(dim=C, mlp_ratio=cfg.mlp_ratio, drop=cfg.ffn_drop, act=cfg.mlp_act)
        self.dp3 = DropPath(cfg.drop_path) if cfg.drop_path > 0 else nn.Identity()

    def _tail_rows(self, rows: Tensor, geom: Geom) -> Tensor:
        s2 = _drop_scale(self.dp2, rows, geom.B)
        rows = self.grid_attn.rows_forward(rows, geom, self.norm2, s2, True)
        s3 = _drop_scale(self.dp3, rows, geom.B)
        return self.mlp.rows_forward(rows, geom, self.norm3, s3, True)


def _make_mbconv(cfg) -> MBConv:
    C = cfg.dim
    return MBConv(in_ch=C, out_ch=C, stride=1,
                  cfg=MBConvConfig(expand_ratio=cfg.mbconv_expand_ratio, se_ratio=cfg.mbconv_se_ratio,
                                   act=cfg.mbconv_act, use_bn=cfg.use_bn, drop_path=0.0))


class OutGridBlock(_GridMlpTail):
    """Outlooker -> MBConv -> Grid-MHSA -> MLP on [B,C,H,W].  (Out_Grid_Block.py:35-107)

    The activation stays in NHWC rows through the whole block; the returned tensor is the logical
    NCHW view of those rows (channels_last strides)."""

    def __init__(self, cfg):
        super().__init__()
        C = cfg.dim
        self.outlook = OutlookerBlock2d(dim=C, num_heads=cfg.outlook_heads, kernel_size=cfg.outlook_kernel, stride=1,
                                        mlp_ratio=cfg.outlook_mlp_ratio, attn_drop=cfg.attn_drop,
                                        proj_drop=cfg.proj_drop, mlp_drop=cfg.ffn_drop, drop_path=cfg.drop_path,
                                        act=cfg.mlp_act)
        self.mbconv = _make_mbconv(cfg)
        self._build_tail(cfg)

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        rows = self.outlook.rows_forward(rows, geom)
        rows = self.mbconv.rows_forward(rows, geom)
        rows = self._tail_rows(rows, geom)
        return OF.from_rows(rows, geom)


class GridOnlyBlock(_GridMlpTail):
    """MBConv -> Grid-MHSA -> MLP (Model B).  (Grid_Only_Block.py:21-73)"""

    def __init__(self, cfg):
        super().__init__()
        self.mbconv = _make_mbconv(cfg)
        self._build_tail(cfg)

    def forward(self, x: Tensor) -> Tensor:
        rows, geom = OF.to_rows(x)
        dt = _compute_dtype(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        rows = self.mbconv.rows_forward(rows, geom)
        rows = self._tail_rows(rows, geom)
        return OF.from_rows(rows, geom)
