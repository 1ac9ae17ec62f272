"""Helpers for the `-m gpu` parity tests: build the product module for a golden case, run it on
cuda:0 through the C-ABI kernels, and hand back outputs/gradients for comparison."""
from __future__ import annotations

import contextlib

import torch

import outlook_grid_vision_transformer_b200 as og
from outlook_grid_vision_transformer_b200 import modules as ogm

DEV = "cuda:0"


def build_module(case):
    kind, st = case["kind"], case["state"]
    if kind == "outlook_attn":
        return og.OutlookAttention2d(st["v.weight"].shape[0], num_heads=case["heads"])
    if kind == "outlooker":
        C = st["attn.v.weight"].shape[0]
        return og.OutlookerBlock2d(C, num_heads=case["heads"], mlp_ratio=st["mlp.fc1.weight"].shape[0] / C,
                                   drop_path=case["drop_path"])
    if kind == "mlp2d":
        C = st["fc1.weight"].shape[1]
        return og.MLP2d(C, mlp_ratio=st["fc1.weight"].shape[0] / C)
    if kind == "mlp":
        C = st["fc1.weight"].shape[1]
        return og.MLP(C, mlp_ratio=st["fc1.weight"].shape[0] / C)
    if kind == "layernorm2d":
        return og.LayerNorm2d(st["ln.weight"].shape[0], eps=1e-6)
    if kind == "mbconv":
        C = st["expand.0.weight"].shape[1]
        return og.MBConv(C, C, 1, og.MBConvConfig(expand_ratio=st["expand.0.weight"].shape[0] / C))
    if kind == "grid_attn":
        return og.GridAttention2D(og.GridAttention2DConfig(mode="grid", dim=st["mhsa.proj.weight"].shape[0],
                                                           num_heads=case["heads"], grid_size=case["grid"]))
    if kind == "outgrid_block":
        return og.OutGridBlock(og.StageCfg(**case["cfg"]))
    if kind == "grid_only_block":
        return og.GridOnlyBlock(og.StageCfg(**case["cfg"]))
    if kind == "model":
        return og.build_model(case["model_cfg"])
    raise KeyError(kind)


@contextlib.contextmanager
def forced_drop_scales(scales):
    """Feed the stochastic-depth scales recorded from the reference run (CPU and CUDA RNG streams differ)."""
    if not scales:
        yield
        return
    queue = [s.to(DEV, torch.float32).contiguous() for s in scales]
    orig = ogm._drop_scale

    def fake(dp, x, batch):
        if not isinstance(dp, ogm.DropPath) or dp.drop_prob == 0.0 or not dp.training:
            return None
        return queue.pop(0)

    ogm._drop_scale = fake
    try:
        yield
    finally:
        ogm._drop_scale = orig


def run_product(case, dtype):
    """-> (y, dx, grads, buffers_after) of the CUDA path for one golden case, compute dtype `dtype`."""
    mod = build_module(case)
    missing = mod.load_state_dict(case["state"], strict=True)
    mod = mod.to(DEV).float().train(case["training"])
    x = case["x"].to(DEV, dtype).clone().requires_grad_(True)
    R = case["R"].to(DEV, torch.float32)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if (dtype == torch.bfloat16 and case["kind"] == "model") else contextlib.nullcontext()
    if case["kind"] == "model":
        x = case["x"].to(DEV, torch.float32).clone().requires_grad_(True)
    with forced_drop_scales(case.get("drop_scales")), ctx:
        y = mod(x)
    (y.float() * R).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in mod.named_parameters() if p.grad is not None}
    bufs = {k: v.detach().cpu() for k, v in mod.state_dict().items() if "running_" in k or "num_batches" in k}
    return y.detach().float().cpu(), x.grad.detach().float().cpu(), grads, bufs
