"""Generate golden input/output vectors by running the UNMODIFIED reference modules.

Run in the authoring container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Writes tests/golden/outgrid_golden.pt: for every case the constructor config, the state_dict, the
input, a fixed cotangent R, the forward output, d(sum(y*R))/dx, every parameter gradient and, for
training-mode BatchNorm, the post-step running statistics.  Everything in float64 so that the
oracle (and through it the CUDA path) is checked against well-conditioned numbers.
Stochastic depth: the per-sample scales the reference drew are recovered from forward hooks on its
DropPath modules (output / input) and stored, because CPU and CUDA generators differ.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

REF = Path(os.environ.get("OGV_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF))

from src.Model_A_OutGridNet import MaxOutNet  # noqa: E402
from src.Model_B_OutGridNet import OutlookerFrontGridNet  # noqa: E402
from src.model.Grid_Only_Block import GridOnlyBlock  # noqa: E402
from src.model.Out_Grid_Block import MLP, OutGridBlock  # noqa: E402
from src.model.Outlook_Block import DropPath, OutlookerBlock2d  # noqa: E402
from src.model.grid_attention import GridAttention2D, GridAttention2DConfig  # noqa: E402
from src.model.grid_partition import grid_partition  # noqa: E402
from src.model.mbc_conv import MBConv, MBConvConfig  # noqa: E402
from src.model.outlook_attention import LayerNorm2d, MLP2d, OutlookAttention2d  # noqa: E402
from src.stage_config import StageCfg  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "outgrid_golden.pt"


def stage_cfg(dim, heads, grid, oheads, drop_path=0.0, expand=2.0, mlp_ratio=2.0):
    return dict(dim=dim, depth=1, num_heads=heads, grid_size=grid, window_size=4, outlook_heads=oheads,
                outlook_kernel=3, outlook_mlp_ratio=2.0, mbconv_expand_ratio=expand, mbconv_se_ratio=0.25,
                mbconv_act="silu", use_bn=True, attn_drop=0.0, proj_drop=0.0, ffn_drop=0.0, drop_path=drop_path,
                mlp_ratio=mlp_ratio, mlp_act="gelu")


def randomize_norms(mod: torch.nn.Module, g: torch.Generator):
    """Default init leaves LN/BN at weight=1,bias=0 and running stats at 0/1; perturb them so a
    wrong gamma/beta/running-stat path cannot hide."""
    for m in mod.modules():
        if isinstance(m, (torch.nn.LayerNorm, torch.nn.BatchNorm2d)):
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g, dtype=m.weight.dtype))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g, dtype=m.bias.dtype))
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g, dtype=m.running_mean.dtype))
                    m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g, dtype=m.running_var.dtype))


def run_case(name, mod, x, training, extra=None):
    g = torch.Generator().manual_seed(1234)
    mod = mod.double()
    randomize_norms(mod, g)
    mod.train(training)
    state_before = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    scales = []
    hooks = []
    for m in mod.modules():
        if isinstance(m, DropPath):
            def hook(module, inp, out, _s=scales):
                if module.drop_prob == 0.0 or not module.training:
                    return
                xin = inp[0]
                B = xin.shape[0]
                keep = 1.0 - module.drop_prob
                alive = (out.reshape(B, -1).abs().sum(dim=1) > 0) | (xin.reshape(B, -1).abs().sum(dim=1) == 0)
                _s.append(torch.where(alive, torch.tensor(1.0 / keep, dtype=torch.float64), torch.tensor(0.0, dtype=torch.float64)))
            hooks.append(m.register_forward_hook(hook))
    x = x.double().clone().requires_grad_(True)
    torch.manual_seed(99)  # DropPath masks
    y = mod(x)
    R = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * R).sum().backward()
    for h in hooks:
        h.remove()
    state_after = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    rec = dict(name=name, training=training, state=state_before, x=x.detach().clone(), R=R, y=y.detach().clone(),
               dx=x.grad.detach().clone(),
               grads={k: p.grad.detach().clone() for k, p in mod.named_parameters() if p.grad is not None},
               buffers_after={k: v for k, v in state_after.items() if "running_" in k or "num_batches" in k},
               drop_scales=scales)
    if extra:
        rec.update(extra)
    print(f"{name:34s} x{tuple(x.shape)} -> y{tuple(y.shape)} |y|={y.abs().mean():.4f} params={len(rec['grads'])} "
          f"drops={len(scales)}")
    return rec


def main():
    torch.manual_seed(0)
    cases = {}
    g = torch.Generator().manual_seed(7)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)

    # --- OutlookAttention2d (NCHW) -------------------------------------------------------------
    torch.manual_seed(1)
    cases["outlook_attn_c16_h4"] = run_case("outlook_attn_c16_h4", OutlookAttention2d(16, num_heads=4), rn(2, 16, 8, 8), True,
                                            dict(kind="outlook_attn", heads=4))
    torch.manual_seed(2)
    cases["outlook_attn_c48_h2"] = run_case("outlook_attn_c48_h2", OutlookAttention2d(48, num_heads=2), rn(1, 48, 6, 5), True,
                                            dict(kind="outlook_attn", heads=2))
    # --- OutlookerBlock2d ------------------------------------------------------------------------
    torch.manual_seed(3)
    cases["outlooker_block_c16"] = run_case("outlooker_block_c16", OutlookerBlock2d(16, num_heads=4, mlp_ratio=2.0),
                                            rn(2, 16, 8, 8), True, dict(kind="outlooker", heads=4, drop_path=0.0))
    torch.manual_seed(4)
    cases["outlooker_block_c16_dp"] = run_case("outlooker_block_c16_dp",
                                               OutlookerBlock2d(16, num_heads=2, mlp_ratio=2.0, drop_path=0.4),
                                               rn(6, 16, 4, 4), True, dict(kind="outlooker", heads=2, drop_path=0.4))
    # --- MLP2d / MLP / LayerNorm2d ---------------------------------------------------------------
    torch.manual_seed(5)
    cases["mlp2d_c16"] = run_case("mlp2d_c16", MLP2d(16, mlp_ratio=2.0), rn(2, 16, 4, 4), True, dict(kind="mlp2d"))
    torch.manual_seed(6)
    cases["mlp_c16"] = run_case("mlp_c16", MLP(16, mlp_ratio=4.0), rn(2, 4, 4, 16), True, dict(kind="mlp"))
    torch.manual_seed(7)
    cases["layernorm2d_c24"] = run_case("layernorm2d_c24", LayerNorm2d(24, eps=1e-6), rn(2, 24, 3, 5), True,
                                        dict(kind="layernorm2d"))
    # --- MBConv ----------------------------------------------------------------------------------
    torch.manual_seed(8)
    cases["mbconv_c16_train"] = run_case("mbconv_c16_train", MBConv(16, 16, 1, MBConvConfig(expand_ratio=2.0)),
                                         rn(3, 16, 8, 8), True, dict(kind="mbconv"))
    torch.manual_seed(9)
    cases["mbconv_c16_eval"] = run_case("mbconv_c16_eval", MBConv(16, 16, 1, MBConvConfig(expand_ratio=2.0)),
                                        rn(3, 16, 8, 8), False, dict(kind="mbconv"))
    torch.manual_seed(10)
    cases["mbconv_c24_train_x4"] = run_case("mbconv_c24_train_x4", MBConv(24, 24, 1, MBConvConfig(expand_ratio=4.0)),
                                            rn(2, 24, 5, 7), True, dict(kind="mbconv"))
    # --- GridAttention2D (BHWC) ------------------------------------------------------------------
    torch.manual_seed(11)
    cases["grid_attn_c16_g2"] = run_case(
        "grid_attn_c16_g2", GridAttention2D(GridAttention2DConfig(mode="grid", dim=16, num_heads=4, grid_size=2)),
        rn(2, 8, 8, 16), True, dict(kind="grid_attn", heads=4, grid=2))
    torch.manual_seed(12)
    cases["grid_attn_c48_g4"] = run_case(
        "grid_attn_c48_g4", GridAttention2D(GridAttention2DConfig(mode="grid", dim=48, num_heads=2, grid_size=4)),
        rn(1, 8, 16, 48), True, dict(kind="grid_attn", heads=2, grid=4))
    # --- OutGridBlock / GridOnlyBlock ------------------------------------------------------------
    torch.manual_seed(13)
    c = stage_cfg(16, 4, 2, 4)
    cases["outgrid_block_c16_train"] = run_case("outgrid_block_c16_train", OutGridBlock(StageCfg(**c)), rn(2, 16, 8, 8),
                                                True, dict(kind="outgrid_block", cfg=c))
    torch.manual_seed(14)
    cases["outgrid_block_c16_eval"] = run_case("outgrid_block_c16_eval", OutGridBlock(StageCfg(**c)), rn(2, 16, 8, 8),
                                               False, dict(kind="outgrid_block", cfg=c))
    torch.manual_seed(15)
    c = stage_cfg(32, 2, 4, 4, drop_path=0.3, expand=4.0, mlp_ratio=4.0)
    cases["outgrid_block_c32_dp"] = run_case("outgrid_block_c32_dp", OutGridBlock(StageCfg(**c)), rn(5, 32, 8, 8), True,
                                             dict(kind="outgrid_block", cfg=c))
    torch.manual_seed(16)
    c = stage_cfg(16, 2, 2, 4)
    cases["grid_only_block_c16"] = run_case("grid_only_block_c16", GridOnlyBlock(StageCfg(**c)), rn(2, 16, 4, 4), True,
                                            dict(kind="grid_only_block", cfg=c))
    # --- whole models (tiny) ---------------------------------------------------------------------
    torch.manual_seed(17)
    mcfg = dict(type="model_a", num_classes=10, in_ch=3, stem_dim=16, dpr_max=0.0,
                stages=[{k: v for k, v in stage_cfg(16, 4, 2, 4).items() if k != "drop_path"},
                        {k: v for k, v in stage_cfg(32, 4, 2, 4).items() if k != "drop_path"}])
    ma = MaxOutNet(num_classes=10, stages=[StageCfg(**s) for s in mcfg["stages"]], stem_dim=16, dpr_max=0.0)
    cases["model_a_tiny"] = run_case("model_a_tiny", ma, rn(2, 3, 8, 8), True, dict(kind="model", model_cfg=mcfg))
    torch.manual_seed(18)
    mcfg_b = dict(mcfg, type="model_b", outlooker_front_depth=1)
    mb = OutlookerFrontGridNet(num_classes=10, stages=[StageCfg(**s) for s in mcfg_b["stages"]], stem_dim=16,
                               outlooker_front_depth=1, dpr_max=0.0)
    cases["model_b_tiny_eval"] = run_case("model_b_tiny_eval", mb, rn(2, 3, 8, 8), False, dict(kind="model", model_cfg=mcfg_b))

    # --- integer index goldens (bit-exact) -------------------------------------------------------
    xi = torch.arange(2 * 8 * 12 * 1, dtype=torch.float64).reshape(2, 8, 12, 1)
    grids, meta = grid_partition(xi, 4)
    cases["index_grid_partition"] = dict(kind="index_grid", B=2, H=8, W=12, g=4, grids=grids.to(torch.long).squeeze(-1).reshape(2 * 16, -1))
    img = torch.arange(1, 1 + 1 * 2 * 5 * 6, dtype=torch.float64).reshape(1, 2, 5, 6)
    unf = torch.nn.functional.unfold(img, kernel_size=3, padding=1, stride=1)  # [1, C*9, H*W]
    cases["index_unfold"] = dict(kind="index_unfold", img=img, unf=unf)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    torch.save(cases, OUT)
    print("wrote", OUT, f"{OUT.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
