"""Shared helpers: run the CPU oracle (oracle/outgrid_oracle.py) on a golden case and compare tensors."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import torch

from oracle import outgrid_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden" / "outgrid_golden.pt"


def load_golden():
    return torch.load(GOLDEN, map_location="cpu", weights_only=False)


def tensor_cases(cases):
    return {k: v for k, v in cases.items() if "state" in v}


def oracle_forward(case, params, x, aux=None):
    """Forward of one golden case through the oracle; x and the output use the case's own layout."""
    kind, training = case["kind"], case["training"]
    P = {("m." + k): v for k, v in params.items()}
    scales = case.get("drop_scales") or None
    if kind == "outlook_attn":
        return O.outlook_attention(x.permute(0, 2, 3, 1), P, "m", case["heads"]).permute(0, 3, 1, 2)
    if kind == "outlooker":
        sc = O._Scales(scales)
        y = O.outlooker_block(x.permute(0, 2, 3, 1), P, "m", case["heads"], "gelu", sc, training and case["drop_path"] > 0)
        return y.permute(0, 3, 1, 2)
    if kind == "mlp2d":
        return O.mlp_rows(x.permute(0, 2, 3, 1), P, "m", "gelu").permute(0, 3, 1, 2)
    if kind == "mlp":
        return O.mlp_rows(x, P, "m", "gelu")
    if kind == "layernorm2d":
        return O.layer_norm(x.permute(0, 2, 3, 1), P["m.ln.weight"], P["m.ln.bias"], 1e-6).permute(0, 3, 1, 2)
    if kind == "mbconv":
        a = {} if aux is not None else None
        y = O.mbconv(x.permute(0, 2, 3, 1), P, "m", "silu", training, a).permute(0, 3, 1, 2)
        if aux is not None:
            aux.update({k[2:]: v for k, v in a.items()})
        return y
    if kind == "grid_attn":
        return O.grid_attention(x, P, "m", case["heads"], case["grid"])
    if kind in ("outgrid_block", "grid_only_block"):
        cfg = SimpleNamespace(**case["cfg"])
        a = {} if aux is not None else None
        fn = O.outgrid_block if kind == "outgrid_block" else O.grid_only_block
        y = fn(x, P, "m", cfg, training, a, scales)
        if aux is not None:
            aux.update({k[2:]: v for k, v in a.items()})
        return y
    if kind == "model":
        return O.model_forward(x, params, case["model_cfg"], training, aux, scales)
    raise KeyError(kind)


def oracle_run(case, dtype=torch.float64):
    """-> (y, dx, grads, aux) from the oracle for one golden case."""
    params = {}
    for k, v in case["state"].items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if "running_" not in k:
                t.requires_grad_(True)
        params[k] = t
    x = case["x"].to(dtype).clone().requires_grad_(True)
    aux = {}
    y = oracle_forward(case, params, x, aux)
    (y * case["R"].to(dtype)).sum().backward()
    grads = {k: p.grad for k, p in params.items() if p.is_floating_point() and p.requires_grad and p.grad is not None}
    return y.detach(), x.grad.detach(), grads, aux


def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / (max|b| + tiny): scale-aware error used with the north_star tolerances."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, rtol, what="", atol=1e-12):
    """|a-b| <= rtol * (|b| + max|b|): elementwise relative tolerance with a tensor-scale floor (so
    near-zero entries are judged against the tensor's own magnitude)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    tol = rtol * (b.abs() + b.abs().max()) + atol  # atol: gradients that are exactly 0 in exact arithmetic
    bad = (a - b).abs() > tol
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol}; "
                             f"max_rel_err={max_rel_err(a, b):.3e}; first at {i}: got {a[tuple(i)]:.6g} want {b[tuple(i)]:.6g}")


def assert_close_rms(a, b, rtol, what="", outlier_frac=0.0, outlier_band=1.0):
    """Plain elementwise criterion for FORWARD outputs: |a-b| <= rtol*|b| + rtol*rms(b)  (what
    torch.testing.assert_close(rtol=rtol, atol=rtol*rms) checks; stricter than the tensor-max band of assert_close).
    bf16 callers may allow a fraction `outlier_frac` of the elements up to `outlier_band` x that band: a block output
    is stored in bf16 (2^-9 relative) behind three BatchNorms, and one element in 1e5 lands a few % past a 2e-2 band."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    rms = float(b.pow(2).mean().sqrt())
    ratio = (a - b).abs() / (rtol * b.abs() + rtol * rms)
    bad = ratio > 1.0
    if bad.any() and not (float(bad.double().mean()) <= outlier_frac and float(ratio.max()) <= outlier_band):
        worst = float(ratio.max())
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol}, atol=rtol*rms={rtol * rms:.3e} "
                             f"(worst {worst:.2f}x the band)")


def normwise_err(a, b) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-3 * b.numel() ** 0.5))


def assert_grad_close_bf16(a, b, rtol, what="", elementwise=True):
    """bf16 criterion for PARAMETER gradients (sums over B*H*W bf16 products with heavy cancellation,
    e.g. BatchNorm gamma): normwise relative error <= rtol, at most 1% of the elements outside the
    elementwise band |a-b| <= rtol*(|b|+max|b|), and no element further than 2.5x that band."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    scale = b.abs().max() + 1e-3
    l2 = float((a - b).norm() / (b.norm() + 1e-3 * b.numel() ** 0.5))
    band = rtol * (b.abs() + scale)
    out = (a - b).abs() > band
    worst = float(((a - b).abs() / band).max())
    assert l2 <= rtol, f"{what}: normwise rel err {l2:.3e} > {rtol}"
    if not elementwise:
        return
    assert out.float().mean() <= 0.01 and worst <= 2.5, (
        f"{what}: {int(out.sum())}/{out.numel()} elements outside the rtol={rtol} band, worst {worst:.2f}x")
