"""CPU ORACLE for the OutGridBlock hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file, and only as the checker / the timed CPU baseline.  The product path
(outlook_grid_vision_transformer_b200) never imports it and has no CPU fallback.

What it is: a from-scratch functional restatement of the reference's algorithm on plain tensors
(a flat `params` dict with the reference's state_dict key names), written as explicit row/shift/
index arithmetic rather than nn.Modules; gradients come from autograd over this restatement.
The arithmetic primitives themselves (matmul, exp, erf, conv for the out-of-scope stem/downsample)
are the third-party dependency the reference also sits on: PyTorch (requirements.txt: torch>=2.0.0,
unpinned; 2.11.0+cu128 installed) -- its CPU kernels, in fp32 or fp64.

Pinning: the reference's own tests hold no golden values for this path, so the oracle
is pinned against outputs of the reference itself, generated in the authoring container by
`oracle/make_golden.py` (imports /root/reference read-only) and committed under tests/golden/;
`tests/test_oracle_golden.py` checks forward values, input gradients, every parameter gradient and
the BatchNorm running statistics.  When /root/reference is present the same test also compares live
against the reference modules on extra shapes.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def act_fn(name: str, x: Tensor) -> Tensor:
    """make_activation: silu | relu | gelu(exact erf).  outlook_attention.py:6-14"""
    name = name.lower()
    if name == "silu":
        return x * torch.sigmoid(x)
    if name == "relu":
        return torch.clamp_min(x, 0)
    if name == "gelu":
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    raise ValueError(f"Unknown activation '{name}'. Use one of: silu|gelu|relu")


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """LayerNorm over the last dim, biased variance.  outlook_attention.py:24-31, Out_Grid_Block.py:69,84"""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def batch_norm(x: Tensor, p: Params, prefix: str, training: bool, aux: Optional[dict], eps: float = 1e-5,
               momentum: float = 0.1) -> Tensor:
    """BatchNorm over all leading dims of a [..., C] tensor; batch statistics in training, running
    statistics in eval; records the running-stat update in `aux`.  mbc_conv.py:61, SURVEY A.3"""
    w, b = p[prefix + ".weight"], p[prefix + ".bias"]
    flat = x.reshape(-1, x.shape[-1])
    if training:
        n = flat.shape[0]
        mean = flat.mean(dim=0)
        var = ((flat - mean) ** 2).mean(dim=0)
        if aux is not None:
            rm, rv = p[prefix + ".running_mean"], p[prefix + ".running_var"]
            unbiased = var * n / max(n - 1, 1)
            aux[prefix + ".running_mean"] = ((1 - momentum) * rm + momentum * mean).detach()
            aux[prefix + ".running_var"] = ((1 - momentum) * rv + momentum * unbiased).detach()
            aux[prefix + ".num_batches_tracked"] = p[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var + eps) * w + b


def pointwise(x: Tensor, p: Params, prefix: str) -> Tensor:
    """1x1 conv / Linear on channels-last data: x @ W^T (+ b).  Conv weights are [out,in,1,1]."""
    w = p[prefix + ".weight"]
    w = w.reshape(w.shape[0], -1)
    y = x @ w.t()
    b = p.get(prefix + ".bias")
    return y if b is None else y + b


def shift2d(x: Tensor, dh: int, dw: int) -> Tensor:
    """y[b,h,w,:] = x[b,h+dh,w+dw,:] with zeros outside the image ([B,H,W,C] layout)."""
    B, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    return xp[:, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W, :]


def drop_path_apply(y: Tensor, scale: Optional[Tensor]) -> Tensor:
    """y * mask / keep with a per-sample scale vector.  Outlook_Block.py:15-22"""
    if scale is None:
        return y
    return y * scale.to(y.dtype).reshape(-1, *([1] * (y.dim() - 1)))


class _Scales:
    """Hands out the per-sample stochastic-depth scale vectors in call order."""

    def __init__(self, scales: Optional[Sequence[Optional[Tensor]]]):
        self.scales = list(scales) if scales is not None else None
        self.i = 0

    def next(self, active: bool) -> Optional[Tensor]:
        if not active or self.scales is None:
            return None
        s = self.scales[self.i]
        self.i += 1
        return s


# ----------------------------------------------------------------------------------------------
# outlook attention  (outlook_attention.py:91-124)
# ----------------------------------------------------------------------------------------------
def outlook_attention(x: Tensor, p: Params, prefix: str, heads: int) -> Tensor:
    """x: [B,H,W,C].  logits channel = head*9 + (ki*3+kj) (:106); softmax over the 9 taps (:107);
    y[b,h,w,c] = sum_t A[b,h,w,head(c),t] * v[b,h+ki-1,w+kj-1,c], zero outside, no renormalisation
    (:110-120, F.unfold ordering c*9 + ki*3 + kj); then the 1x1 proj (:122)."""
    B, H, W, C = x.shape
    hd = C // heads
    logits = pointwise(x, p, prefix + ".attn").reshape(B, H, W, heads, 9)
    A = torch.softmax(logits, dim=-1)
    v = pointwise(x, p, prefix + ".v")
    y = torch.zeros_like(v).reshape(B, H, W, heads, hd)
    for t in range(9):
        ki, kj = divmod(t, 3)
        y = y + A[..., t].unsqueeze(-1) * shift2d(v, ki - 1, kj - 1).reshape(B, H, W, heads, hd)
    return pointwise(y.reshape(B, H, W, C), p, prefix + ".proj")


def mlp_rows(x: Tensor, p: Params, prefix: str, act: str) -> Tensor:
    """fc1 -> act -> fc2 on the channel dim.  outlook_attention.py:43-49 (MLP2d), Out_Grid_Block.py:24-32 (MLP)"""
    return pointwise(act_fn(act, pointwise(x, p, prefix + ".fc1")), p, prefix + ".fc2")


def outlooker_block(x: Tensor, p: Params, prefix: str, heads: int, act: str, scales: _Scales, dp_active: bool,
                    eps: float = 1e-6) -> Tensor:
    """x + dp1(attn(norm1 x)); x + dp2(mlp(norm2 x)).  Outlook_Block.py:61-64"""
    y = outlook_attention(layer_norm(x, p[prefix + ".norm1.ln.weight"], p[prefix + ".norm1.ln.bias"], eps), p,
                          prefix + ".attn", heads)
    x = x + drop_path_apply(y, scales.next(dp_active))
    y = mlp_rows(layer_norm(x, p[prefix + ".norm2.ln.weight"], p[prefix + ".norm2.ln.bias"], eps), p,
                 prefix + ".mlp", act)
    return x + drop_path_apply(y, scales.next(dp_active))


# ----------------------------------------------------------------------------------------------
# MBConv  (mbc_conv.py:90-98)
# ----------------------------------------------------------------------------------------------
def mbconv(x: Tensor, p: Params, prefix: str, act: str, training: bool, aux: Optional[dict]) -> Tensor:
    """expand(1x1,BN,act) -> depthwise(3x3,BN,act) -> SE -> project(1x1,BN) -> x + out."""
    B, H, W, C = x.shape
    e = act_fn(act, batch_norm(pointwise(x, p, prefix + ".expand.0"), p, prefix + ".expand.1", training, aux))
    wdw = p[prefix + ".depthwise.0.weight"]  # [Cm,1,3,3]
    d = torch.zeros_like(e)
    for t in range(9):
        ki, kj = divmod(t, 3)
        d = d + shift2d(e, ki - 1, kj - 1) * wdw[:, 0, ki, kj]
    d = act_fn(act, batch_norm(d, p, prefix + ".depthwise.1", training, aux))
    s = d.mean(dim=(1, 2))                                                    # squeeze, mbc_conv.py:22-23
    s = act_fn(act, pointwise(s, p, prefix + ".se.fc1"))
    s = torch.sigmoid(pointwise(s, p, prefix + ".se.fc2"))
    d = d * s[:, None, None, :]                                               # excite, :27
    o = batch_norm(pointwise(d, p, prefix + ".project.0"), p, prefix + ".project.1", training, aux)
    return x + o


# ----------------------------------------------------------------------------------------------
# grid attention  (grid_partition.py:13-15, grid_attention.py:62-89)
# ----------------------------------------------------------------------------------------------
def grid_index_loops(B: int, H: int, W: int, g: int) -> Tensor:
    """rows[bg, n] = flat NHWC position of token n of group bg; group = b*g*g + (h%g)*g + (w%g),
    token = (h//g)*Wg + (w//g).  Plain integer loops: the bit-exact statement of grid_partition
    (small cases only)."""
    if g <= 0:
        raise ValueError("grid_size must be > 0")
    if H % g or W % g:
        raise ValueError(f"H and W must be divisible by grid_size. Got H={H}, W={W}, g={g}")
    Hg, Wg = H // g, W // g
    idx = torch.empty((B * g * g, Hg * Wg), dtype=torch.long)
    for b in range(B):
        for gi in range(g):
            for gj in range(g):
                bg = (b * g + gi) * g + gj
                for hg in range(Hg):
                    for wg in range(Wg):
                        idx[bg, hg * Wg + wg] = (b * H + hg * g + gi) * W + wg * g + gj
    return idx


def grid_index(B: int, H: int, W: int, g: int) -> Tensor:
    """Vectorised grid_index_loops (same integers; tests check they agree)."""
    if g <= 0:
        raise ValueError("grid_size must be > 0")
    if H % g or W % g:
        raise ValueError(f"H and W must be divisible by grid_size. Got H={H}, W={W}, g={g}")
    Hg, Wg = H // g, W // g
    b = torch.arange(B).view(B, 1, 1, 1, 1)
    gi = torch.arange(g).view(1, g, 1, 1, 1)
    gj = torch.arange(g).view(1, 1, g, 1, 1)
    hg = torch.arange(Hg).view(1, 1, 1, Hg, 1)
    wg = torch.arange(Wg).view(1, 1, 1, 1, Wg)
    idx = (b * H + hg * g + gi) * W + wg * g + gj
    return idx.reshape(B * g * g, Hg * Wg)


def grid_attention(x: Tensor, p: Params, prefix: str, heads: int, g: int, return_attn: bool = False):
    """x: [B,H,W,C] -> same.  qkv channel = which*C + head*hd + d (:70); softmax((q k^T) hd^-0.5) v; proj."""
    B, H, W, C = x.shape
    hd = C // heads
    idx = grid_index(B, H, W, g)
    Bg, N = idx.shape
    tok = x.reshape(B * H * W, C)[idx.reshape(-1)].reshape(Bg, N, C)
    qkv = pointwise(tok, p, prefix + ".mhsa.qkv").reshape(Bg, N, 3, heads, hd)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # [Bg, heads, N, hd]
    attn = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(Bg, N, C)
    out = pointwise(out, p, prefix + ".mhsa.proj")
    y = torch.zeros(B * H * W, C, dtype=out.dtype).index_copy(0, idx.reshape(-1), out.reshape(Bg * N, C))
    y = y.reshape(B, H, W, C)
    return (y, attn) if return_attn else y


# ----------------------------------------------------------------------------------------------
# blocks  (Out_Grid_Block.py:88-107, Grid_Only_Block.py:58-73)
# ----------------------------------------------------------------------------------------------
def _grid_mlp_tail(x: Tensor, p: Params, prefix: str, cfg, scales: _Scales, dp_active: bool) -> Tensor:
    y = grid_attention(layer_norm(x, p[prefix + ".norm2.weight"], p[prefix + ".norm2.bias"], 1e-5), p,
                       prefix + ".grid_attn", cfg.num_heads, cfg.grid_size)
    x = x + drop_path_apply(y, scales.next(dp_active))
    y = mlp_rows(layer_norm(x, p[prefix + ".norm3.weight"], p[prefix + ".norm3.bias"], 1e-5), p, prefix + ".mlp",
                 cfg.mlp_act)
    return x + drop_path_apply(y, scales.next(dp_active))


def outgrid_block(x_nchw: Tensor, p: Params, prefix: str, cfg, training: bool, aux: Optional[dict] = None,
                  drop_scales: Optional[Sequence[Optional[Tensor]]] = None) -> Tensor:
    """NCHW in / NCHW out; `drop_scales` = the four per-sample scale vectors in the order
    outlook.dp1, outlook.dp2, dp2, dp3 (only consumed when training and cfg.drop_path > 0)."""
    scales = drop_scales if isinstance(drop_scales, _Scales) else _Scales(drop_scales)
    dp_active = training and cfg.drop_path > 0
    x = x_nchw.permute(0, 2, 3, 1)
    x = outlooker_block(x, p, prefix + ".outlook", cfg.outlook_heads, cfg.mlp_act, scales, dp_active)
    x = mbconv(x, p, prefix + ".mbconv", cfg.mbconv_act, training, aux)
    x = _grid_mlp_tail(x, p, prefix, cfg, scales, dp_active)
    return x.permute(0, 3, 1, 2)


def grid_only_block(x_nchw: Tensor, p: Params, prefix: str, cfg, training: bool, aux: Optional[dict] = None,
                    drop_scales=None) -> Tensor:
    scales = drop_scales if isinstance(drop_scales, _Scales) else _Scales(drop_scales)
    dp_active = training and cfg.drop_path > 0
    x = x_nchw.permute(0, 2, 3, 1)
    x = mbconv(x, p, prefix + ".mbconv", cfg.mbconv_act, training, aux)
    x = _grid_mlp_tail(x, p, prefix, cfg, scales, dp_active)
    return x.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# whole models (the callers; stem / downsample / head are outside the hot path and use F.conv2d)
# ----------------------------------------------------------------------------------------------
def _conv_bn_act(x: Tensor, p: Params, conv: str, bn: str, stride: int, training: bool, aux) -> Tensor:
    y = F.conv2d(x, p[conv + ".weight"], p.get(conv + ".bias"), stride=stride, padding=1)
    y = batch_norm(y.permute(0, 2, 3, 1), p, bn, training, aux).permute(0, 3, 1, 2)
    return act_fn("silu", y)


def make_dpr(total_blocks: int, dpr_max: float) -> List[float]:
    """stem_head.py:17-20"""
    if total_blocks <= 1:
        return [dpr_max]
    return [dpr_max * i / (total_blocks - 1) for i in range(total_blocks)]


def model_forward(x: Tensor, p: Params, model_cfg: dict, training: bool, aux: Optional[dict] = None,
                  drop_scales: Optional[Sequence[Optional[Tensor]]] = None) -> Tensor:
    """MaxOutNet.forward (Model_A_OutGridNet.py:55-67) / OutlookerFrontGridNet.forward
    (Model_B_OutGridNet.py:76-100) on a params dict with the reference's state_dict keys."""
    from types import SimpleNamespace

    defaults = dict(window_size=8, outlook_heads=6, outlook_kernel=3, outlook_mlp_ratio=2.0, mbconv_expand_ratio=4.0,
                    mbconv_se_ratio=0.25, mbconv_act="silu", use_bn=True, attn_drop=0.0, proj_drop=0.0, ffn_drop=0.0,
                    drop_path=0.0, mlp_ratio=4.0, mlp_act="gelu")
    stages = [SimpleNamespace(**{**defaults, **s}) for s in model_cfg["stages"]]
    mtype = str(model_cfg.get("type", "model_a")).lower()
    is_b = mtype in ("b", "model_b", "outlooker_front", "front")
    front = int(model_cfg.get("outlooker_front_depth", 2)) if is_b else 0
    dprs = make_dpr(front + sum(s.depth for s in stages), float(model_cfg.get("dpr_max", 0.1)))
    scales = _Scales(drop_scales)
    x = _conv_bn_act(x, p, "stem.stem.0", "stem.stem.1", 1, training, aux)
    if "proj_in.weight" in p:
        x = pointwise(x.permute(0, 2, 3, 1), p, "proj_in").permute(0, 3, 1, 2)
    idx = 0
    for i in range(front):
        c = stages[0]
        xb = outlooker_block(x.permute(0, 2, 3, 1), p, f"front.{i}", c.outlook_heads, c.mlp_act, scales,
                             training and dprs[idx] > 0)
        x = xb.permute(0, 3, 1, 2)
        idx += 1
    for si, s in enumerate(stages):
        for bi in range(s.depth):
            bcfg = SimpleNamespace(**{**s.__dict__, "drop_path": dprs[idx]})
            fn = grid_only_block if is_b else outgrid_block
            x = fn(x, p, f"stages.{si}.{bi}", bcfg, training, aux, scales)
            idx += 1
        if si < len(stages) - 1:
            x = _conv_bn_act(x, p, f"downs.{si}.op.0", f"downs.{si}.op.1", 2, training, aux)
    x = batch_norm(x.permute(0, 2, 3, 1), p, "head_norm", training, aux)
    return pointwise(x.mean(dim=(1, 2)), p, "classifier")


def init_params(model_cfg: dict, seed: int = 0, dtype=torch.float32) -> Params:
    """Random parameters with the reference's shapes/keys (PyTorch default-init statistics), built
    without the reference or the product package: used by the CPU baseline timer."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, out_f, in_f, bias=True, conv=False, groups_k=None):
        fan_in = in_f if groups_k is None else groups_k
        bound = 1.0 / math.sqrt(fan_in)
        shape = (out_f, in_f, 1, 1) if conv else (out_f, in_f)
        if groups_k is not None:
            shape = (out_f, 1, 3, 3)
        p[name + ".weight"] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
        if bias:
            p[name + ".bias"] = (torch.rand(out_f, generator=g, dtype=dtype) * 2 - 1) * bound

    def norm(name, c, bn=False):
        p[name + ".weight"] = torch.ones(c, dtype=dtype)
        p[name + ".bias"] = torch.zeros(c, dtype=dtype)
        if bn:
            p[name + ".running_mean"] = torch.zeros(c, dtype=dtype)
            p[name + ".running_var"] = torch.ones(c, dtype=dtype)
            p[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def outlooker(pre, C, heads):
        norm(pre + ".norm1.ln", C); lin(pre + ".attn.attn", heads * 9, C, conv=True); lin(pre + ".attn.v", C, C, conv=True)
        lin(pre + ".attn.proj", C, C, conv=True); norm(pre + ".norm2.ln", C)
        lin(pre + ".mlp.fc1", 2 * C, C, conv=True); lin(pre + ".mlp.fc2", C, 2 * C, conv=True)

    def tail(pre, C, with_outlook, s):
        if with_outlook:
            outlooker(pre + ".outlook", C, s.get("outlook_heads", 6))
        Cm = 4 * C
        lin(pre + ".mbconv.expand.0", Cm, C, bias=False, conv=True); norm(pre + ".mbconv.expand.1", Cm, bn=True)
        lin(pre + ".mbconv.depthwise.0", Cm, Cm, bias=False, groups_k=9); norm(pre + ".mbconv.depthwise.1", Cm, bn=True)
        lin(pre + ".mbconv.se.fc1", C, Cm, conv=True); lin(pre + ".mbconv.se.fc2", Cm, C, conv=True)
        lin(pre + ".mbconv.project.0", C, Cm, bias=False, conv=True); norm(pre + ".mbconv.project.1", C, bn=True)
        norm(pre + ".norm2", C); lin(pre + ".grid_attn.mhsa.qkv", 3 * C, C); lin(pre + ".grid_attn.mhsa.proj", C, C)
        norm(pre + ".norm3", C); lin(pre + ".mlp.fc1", 4 * C, C); lin(pre + ".mlp.fc2", C, 4 * C)

    stages = model_cfg["stages"]
    stem = int(model_cfg.get("stem_dim", 64))
    in_ch = int(model_cfg.get("in_ch", 3))
    bound = 1.0 / math.sqrt(in_ch * 9)
    p["stem.stem.0.weight"] = (torch.rand((stem, in_ch, 3, 3), generator=g, dtype=dtype) * 2 - 1) * bound
    norm("stem.stem.1", stem, bn=True)
    if stem != stages[0]["dim"]:
        lin("proj_in", stages[0]["dim"], stem, conv=True)
    is_b = str(model_cfg.get("type", "model_a")).lower() in ("b", "model_b", "outlooker_front", "front")
    if is_b:
        for i in range(int(model_cfg.get("outlooker_front_depth", 2))):
            outlooker(f"front.{i}", stages[0]["dim"], stages[0].get("outlook_heads", 6))
    for si, s in enumerate(stages):
        for bi in range(s["depth"]):
            tail(f"stages.{si}.{bi}", s["dim"], not is_b, s)
        if si < len(stages) - 1:
            cin, cout = s["dim"], stages[si + 1]["dim"]
            b2 = 1.0 / math.sqrt(cin * 9)
            p[f"downs.{si}.op.0.weight"] = (torch.rand((cout, cin, 3, 3), generator=g, dtype=dtype) * 2 - 1) * b2
            norm(f"downs.{si}.op.1", cout, bn=True)
    norm("head_norm", stages[-1]["dim"], bn=True)
    lin("classifier", int(model_cfg.get("num_classes", 100)), stages[-1]["dim"])
    return p
