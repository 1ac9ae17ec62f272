"""The callers the hot path drops into: Model A (MaxOutNet, src/Model_A_OutGridNet.py:9-67) and
Model B (OutlookerFrontGridNet, src/Model_B_OutGridNet.py:11-100), re-stated so the framework is
self-contained on a box without the reference checkout.  Constructor signatures, attribute names
and state_dict keys match the reference, so its checkpoints load with strict=True.

SURVEY section 8 row (f2): the conv -> BatchNorm -> act units of the stem and the Downsample layers run on this
package's kernels (functional.ConvBnActFn): the stem convolution as patches x tcgen05 GEMM, the 3x3 stride-2 Downsample
convolution as an implicit GEMM (5-D TMA boxes of the image feed the tcgen05 pipeline; input gradient through a
col2im gather), batch statistics, normalise + activation and their backward on the streaming kernels.  The head
(BatchNorm on [B, C, 4, 4], global average pool, classifier) is functional.HeadFn: statistics, normalise + pool in one
pass, the classifier and its gradients on the GEMM engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Literal

import torch
import torch.nn as nn

from . import functional as OF
from .config import StageCfg, build_stages
from .modules import GridOnlyBlock, OutGridBlock, OutlookerBlock2d, _act_name, _compute_dtype, _direct, make_activation

DownsampleType = Literal["conv", "pool"]


def _conv_bn_act(seq: nn.Sequential, x: torch.Tensor, owner: nn.Module = None) -> torch.Tensor:
    """Run a (Conv2d, BatchNorm2d, act) Sequential: fused BatchNorm + activation kernels on CUDA; the plain module chain
    when a hook sits on one of the three (it must see the tensors the reference would hand it), when the layout is not
    the fused one (no BatchNorm, conv bias, channel count not a multiple of 8), or off the GPU (construction-time shape
    checks on CPU; the blocks themselves refuse CPU tensors)."""
    conv, bn, act = seq[-3], seq[-2], seq[-1]
    fused = (x.is_cuda and isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and conv.bias is None
             and conv.groups == 1 and conv.dilation == (1, 1) and conv.padding_mode == "zeros"
             and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
             and conv.out_channels % 8 == 0 and bn.affine and bn.track_running_stats
             and not any(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks for m in (conv, bn, act)))
    if not fused:
        return seq(x)
    for m in seq[:-3]:  # the AvgPool2d of the "pool" kind
        x = m(x)
    training = bn.training
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    if training and bn.momentum is None:  # cumulative moving average (nn.BatchNorm2d semantics)
        momentum = 1.0 / float(bn.num_batches_tracked)
    else:
        momentum = bn.momentum if bn.momentum is not None else 0.0
    return OF.conv_bn_act(x, conv.weight, bn.weight, bn.bias, running=(bn.running_mean, bn.running_var),
                          stride=conv.stride[0], padding=conv.padding[0], act=_act_name(act), eps=bn.eps,
                          momentum=momentum, training=training, dtype=_compute_dtype(x),
                          direct=owner is not None and _direct(owner))


def make_dpr(total_blocks: int, dpr_max: float) -> List[float]:
    """Linear stochastic-depth schedule over all blocks.  (stem_head.py:17-20)"""
    if total_blocks <= 1:
        return [dpr_max]
    return [dpr_max * i / (total_blocks - 1) for i in range(total_blocks)]


class ConvStem(nn.Module):
    """3x3 stride-1 conv + BN + act.  (stem_head.py:23-32)"""

    def __init__(self, in_ch: int, out_ch: int, act: str = "silu", use_bn: bool = True):
        super().__init__()
        norm = nn.BatchNorm2d(out_ch) if use_bn else nn.Identity()
        self.stem = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=1, padding=1, bias=not use_bn), norm,
                                  make_activation(act))

    def forward(self, x):
        return _conv_bn_act(self.stem, x, self)


@dataclass(frozen=True)
class DownsampleConfig:
    kind: DownsampleType = "conv"
    act: str = "silu"
    use_bn: bool = True


class Downsample(nn.Module):
    """conv: 3x3 s2 + BN + act; pool: avgpool 2x2 + 1x1 + BN + act.  (downsampling.py:28-65)"""

    def __init__(self, in_ch: int, out_ch: int, cfg: DownsampleConfig = DownsampleConfig()):
        super().__init__()
        if in_ch <= 0 or out_ch <= 0:
            raise ValueError("in_ch and out_ch must be > 0")
        self.in_ch, self.out_ch, self.kind = in_ch, out_ch, cfg.kind
        norm = nn.BatchNorm2d(out_ch) if cfg.use_bn else nn.Identity()
        act = make_activation(cfg.act)
        if cfg.kind == "conv":
            self.op = nn.Sequential(
                nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=2, padding=1, bias=not cfg.use_bn), norm, act)
        elif cfg.kind == "pool":
            self.op = nn.Sequential(
                nn.AvgPool2d(kernel_size=2, stride=2),
                nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=1, padding=0, bias=not cfg.use_bn), norm, act)
        else:
            raise ValueError("cfg.kind must be 'conv' or 'pool'")

    def forward(self, x):
        return _conv_bn_act(self.op, x, self)


def _with_drop_path(scfg: StageCfg, p: float) -> StageCfg:
    return StageCfg(**{**scfg.__dict__, "drop_path": p})


class _Backbone(nn.Module):
    def _build_common(self, num_classes, stages, in_ch, stem_dim):
        assert len(stages) >= 1
        self.stem = ConvStem(in_ch, stem_dim, act="silu", use_bn=True)
        self.proj_in = nn.Identity()
        if stem_dim != stages[0].dim:
            self.proj_in = nn.Conv2d(stem_dim, stages[0].dim, kernel_size=1, bias=True)

    def _build_head(self, stages, num_classes):
        self.head_norm = nn.BatchNorm2d(stages[-1].dim)
        self.classifier = nn.Linear(stages[-1].dim, num_classes)

    def _run_stages(self, x):
        for si, blocks in enumerate(self.stages):
            for blk in blocks:
                x = blk(x)
            if si < len(self.downs):
                x = self.downs[si](x)
        return x

    def _head(self, x):
        bn, fc = self.head_norm, self.classifier
        fused = (x.is_cuda and isinstance(bn, nn.BatchNorm2d) and isinstance(fc, nn.Linear) and bn.affine
                 and bn.track_running_stats and x.shape[1] % 8 == 0
                 and not any(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks for m in (bn, fc)))
        if not fused:
            x = bn(x)
            return fc(x.mean(dim=(2, 3)))
        training = bn.training
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        if training and bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
        else:
            momentum = bn.momentum if bn.momentum is not None else 0.0
        dt = _compute_dtype(x)
        rows, geom = OF.to_rows(x)
        rows = rows.to(dt) if rows.dtype != dt else rows
        return OF.head(rows, bn.weight, bn.bias, fc.weight, fc.bias, B=geom.B, HW=geom.H * geom.W,
                       running=(bn.running_mean, bn.running_var), eps=bn.eps, momentum=momentum, training=training,
                       dtype=dt, direct=_direct(self))


class MaxOutNet(_Backbone):
    """Model A: stem -> [OutGridBlock x depth -> Downsample] x S -> BN -> GAP -> Linear."""

    def __init__(self, num_classes: int, stages: List[StageCfg], in_ch: int = 3, stem_dim: int = 64,
                 dpr_max: float = 0.1, down_cfg: DownsampleConfig = DownsampleConfig(kind="conv", act="silu", use_bn=True)):
        super().__init__()
        self._build_common(num_classes, stages, in_ch, stem_dim)
        dprs = make_dpr(sum(s.depth for s in stages), dpr_max)
        idx = 0
        self.stages = nn.ModuleList()
        self.downs = nn.ModuleList()
        for si, scfg in enumerate(stages):
            blocks = nn.ModuleList()
            for _ in range(scfg.depth):
                blocks.append(OutGridBlock(_with_drop_path(scfg, dprs[idx])))
                idx += 1
            self.stages.append(blocks)
            if si < len(stages) - 1:
                self.downs.append(Downsample(scfg.dim, stages[si + 1].dim, cfg=down_cfg))
        self._build_head(stages, num_classes)

    def forward(self, x):
        x = self.proj_in(self.stem(x))
        return self._head(self._run_stages(x))


class OutlookerFrontGridNet(_Backbone):
    """Model B: stem -> OutlookerBlock2d x L -> [GridOnlyBlock x depth -> Downsample] x S -> head."""

    def __init__(self, num_classes: int, stages: List[StageCfg], in_ch: int = 3, stem_dim: int = 64,
                 outlooker_front_depth: int = 2, dpr_max: float = 0.1,
                 down_cfg: DownsampleConfig = DownsampleConfig(kind="conv", act="silu", use_bn=True)):
        super().__init__()
        self._build_common(num_classes, stages, in_ch, stem_dim)
        dprs = make_dpr(outlooker_front_depth + sum(s.depth for s in stages), dpr_max)
        idx = 0
        c = stages[0]
        self.front = nn.ModuleList()
        for _ in range(outlooker_front_depth):
            self.front.append(OutlookerBlock2d(dim=c.dim, num_heads=c.outlook_heads, kernel_size=c.outlook_kernel,
                                               stride=1, mlp_ratio=c.outlook_mlp_ratio, attn_drop=c.attn_drop,
                                               proj_drop=c.proj_drop, mlp_drop=c.ffn_drop, drop_path=dprs[idx],
                                               act=c.mlp_act))
            idx += 1
        self.stages = nn.ModuleList()
        self.downs = nn.ModuleList()
        for si, scfg in enumerate(stages):
            blocks = nn.ModuleList()
            for _ in range(scfg.depth):
                blocks.append(GridOnlyBlock(_with_drop_path(scfg, dprs[idx])))
                idx += 1
            self.stages.append(blocks)
            if si < len(stages) - 1:
                self.downs.append(Downsample(scfg.dim, stages[si + 1].dim, cfg=down_cfg))
        self._build_head(stages, num_classes)

    def forward(self, x):
        x = self.proj_in(self.stem(x))
        for blk in self.front:
            x = blk(x)
        return self._head(self._run_stages(x))


def build_model(model_cfg: dict) -> nn.Module:
    """YAML `model:` section -> model (scripts/train.py:33-60)."""
    model_type = str(model_cfg.get("type", "model_a")).lower()
    stages = build_stages(model_cfg.get("stages", []))
    if not stages:
        raise ValueError("model.stages must have at least one stage config")
    down_cfg = DownsampleConfig(**model_cfg.get("downsample", {}))
    common = dict(num_classes=int(model_cfg.get("num_classes", 100)), stages=stages,
                  in_ch=int(model_cfg.get("in_ch", 3)), stem_dim=int(model_cfg.get("stem_dim", 64)),
                  dpr_max=float(model_cfg.get("dpr_max", 0.1)), down_cfg=down_cfg)
    if model_type in ("a", "model_a", "maxout", "outgrid"):
        return MaxOutNet(**common)
    if model_type in ("b", "model_b", "outlooker_front", "front"):
        return OutlookerFrontGridNet(outlooker_front_depth=int(model_cfg.get("outlooker_front_depth", 2)), **common)
    raise ValueError(f"Unknown model.type '{model_type}'. Use 'model_a' (MaxOutNet) or 'model_b' (OutlookerFrontGridNet)")
