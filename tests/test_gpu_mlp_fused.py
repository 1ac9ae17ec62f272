"""-m gpu: the fused MLP kernel (ogv_mlp_fwd: fc1 -> activation -> fc2 -> scale -> residual, hidden tile on chip)
against a plain PyTorch fp32 restatement of MLP2d / MLP (outlook_attention.py:43-49, Out_Grid_Block.py:24-32) on the
same bf16 inputs, and against the unfused ogv_gemm route it replaces."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ACTS = {"gelu": lambda t: torch.nn.functional.gelu(t), "silu": torch.nn.functional.silu, "relu": torch.relu}


def _case(M, C, Hd, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, C, generator=g)
    w1 = torch.randn(Hd, C, generator=g) / C ** 0.5
    b1 = 0.2 * torch.randn(Hd, generator=g)
    w2 = torch.randn(C, Hd, generator=g) / Hd ** 0.5
    b2 = 0.2 * torch.randn(C, generator=g)
    res = torch.randn(M, C, generator=g)
    bf = lambda t: t.to(DEV, torch.bfloat16)  # noqa: E731
    return bf(x), bf(w1), b1.to(DEV), bf(w2), b2.to(DEV), bf(res)


def _reference(x, w1, b1, w2, b2, res, scale, rows_per_scale, act):
    h = ACTS[act](x.float() @ w1.float().t() + b1)
    h = h.to(torch.bfloat16).float()  # the hidden tile is a bf16 tensor-core operand, as under autocast
    y = h @ w2.float().t() + b2
    if scale is not None:
        y = y * scale.repeat_interleave(rows_per_scale)[:, None]
    if res is not None:
        y = y + res.float()
    return y


@pytest.mark.parametrize("act", ["gelu", "silu", "relu"])
@pytest.mark.parametrize("M,C,Hd", [(128, 64, 128), (256, 64, 256), (1000, 64, 256), (4096 + 40, 64, 128), (128, 128, 256),
                                    (640, 128, 512), (1000, 128, 512), (148 * 128 * 2 + 24, 64, 256)])
def test_mlp_fwd_matches_fp32_restatement(M, C, Hd, act):
    from outlook_grid_vision_transformer_b200 import ops
    assert ops.mlp_fused_supported(C, Hd, torch.bfloat16)
    x, w1, b1, w2, b2, res = _case(M, C, Hd, seed=M + C + Hd)
    for with_res, with_scale in ((True, True), (False, False), (True, False), (False, True)):
        rps = 8
        scale = None
        if with_scale:
            n = (M + rps - 1) // rps
            scale = ((torch.arange(n, device=DEV) % 3 != 0).float() / 0.75).contiguous()
        y = ops.mlp_fwd(x, w1, b1, w2, b2, act=act, residual=res if with_res else None, row_scale=scale, rows_per_scale=rps)
        torch.cuda.synchronize()
        want = _reference(x, w1, b1, w2, b2, res if with_res else None, scale, rps, act)[:M]
        err = (y.float() - want).abs()
        tol = 2e-2 * want.abs() + 2e-2 * float(want.pow(2).mean().sqrt())
        assert bool((err <= tol).all()), f"res={with_res} scale={with_scale}: worst {float((err / tol).max()):.2f}x the band"
        # and it agrees with the two-GEMM route it replaces to bf16 rounding
        h = torch.empty(M, Hd, device=DEV, dtype=torch.bfloat16)
        ops.gemm(x, w1, h, bias=b1, act=act)
        y2 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
        ops.gemm(h, w2, y2, bias=b2, row_scale=scale, rows_per_scale=rps, residual=res if with_res else None)
        torch.testing.assert_close(y.float(), y2.float(), rtol=2e-2, atol=2e-2 * float(want.abs().mean()))


def test_mlp_fwd_unsupported_shapes_are_reported():
    from outlook_grid_vision_transformer_b200 import ops
    assert not ops.mlp_fused_supported(48, 192, torch.bfloat16)
    assert not ops.mlp_fused_supported(64, 192, torch.bfloat16)
    assert not ops.mlp_fused_supported(64, 256, torch.float32)
    x, w1, b1, w2, b2, _ = _case(128, 64, 256)
    with pytest.raises((NotImplementedError, ValueError)):
        ops.mlp_fwd(x[:, :48].contiguous(), w1[:192, :48].contiguous(), b1[:192].contiguous(), w2[:48, :192].contiguous(),
                    b2[:48].contiguous(), act="gelu")


@pytest.mark.parametrize("act", ["gelu", "silu", "relu"])
@pytest.mark.parametrize("M,C,Hd", [(128, 64, 128), (1000, 64, 256), (4096 + 40, 64, 256), (128, 128, 256), (1000, 128, 512),
                                    (148 * 128 + 24, 128, 512), (148 * 128 * 2 + 24, 64, 256)])
def test_mlp_bwd_matches_autograd_restatement(M, C, Hd, act):
    """ogv_mlp_bwd (hidden activation recomputed on chip) against autograd over the fp32 restatement."""
    from outlook_grid_vision_transformer_b200 import ops
    x, w1, b1, w2, b2, dy = _case(M, C, Hd, seed=7 * M + C)
    for with_scale in (False, True):
        rps = 8
        scale = None
        if with_scale:
            n = (M + rps - 1) // rps
            scale = ((torch.arange(n, device=DEV) % 3 != 0).float() / 0.75).contiguous()
        dxn, dz, hs = ops.mlp_bwd(x, dy, w1, w2.t().contiguous(), w1.t().contiguous(), b1, act=act, row_scale=scale,
                                  rows_per_scale=rps)
        torch.cuda.synchronize()
        s = scale.repeat_interleave(rps)[:M, None] if scale is not None else 1.0
        z = (x.float() @ w1.float().t() + b1).requires_grad_(True)
        h = ACTS[act](z)
        dh = (dy.float() @ w2.float())
        (dact,) = torch.autograd.grad(h, z, torch.ones_like(h))
        dz_want = dh * dact * s
        hs_want = h.detach() * s
        dxn_want = dz_want.to(torch.bfloat16).float() @ w1.float()

        def check(got, want, what):
            err = (got.float() - want).abs()
            tol = 2e-2 * want.abs() + 2e-2 * float(want.pow(2).mean().sqrt())
            assert bool((err <= tol).all()), f"{what} (scale={with_scale}): worst {float((err / tol).max()):.2f}x the band"

        check(dz, dz_want, "dz")
        check(hs, hs_want, "hs")
        check(dxn, dxn_want, "dxn")


@pytest.mark.parametrize("C,H,ratio,kind", [(64, 32, 4.0, "mlp"), (64, 16, 2.0, "mlp2d"), (128, 16, 4.0, "mlp"), (128, 8, 2.0, "mlp2d")])
def test_fused_and_unfused_mlp_branch_agree(C, H, ratio, kind):
    """The whole branch (LN -> MLP -> DropPath -> +residual), forward and every gradient: fused kernels vs the
    two-GEMM route (functional.MLP_FUSED off), same inputs."""
    import outlook_grid_vision_transformer_b200 as og
    from outlook_grid_vision_transformer_b200 import functional as OF
    from outlook_grid_vision_transformer_b200 import modules as M_
    torch.manual_seed(C + H)
    B = 3
    mod = (og.MLP(C, mlp_ratio=ratio) if kind == "mlp" else og.MLP2d(C, mlp_ratio=ratio)).to(DEV).train()
    ln = torch.nn.LayerNorm(C, eps=1e-5).to(DEV)
    with torch.no_grad():
        ln.weight.add_(0.1 * torch.randn_like(ln.weight))
        ln.bias.add_(0.1 * torch.randn_like(ln.bias))
    x = torch.randn(B * H * H, C, device=DEV).bfloat16()
    R = torch.randn(B * H * H, C, device=DEV)
    scale = torch.tensor([1 / 0.75, 0.0, 1 / 0.75], device=DEV)
    outs = []
    for fused in (True, False):
        OF.MLP_FUSED = fused
        try:
            for p in list(mod.parameters()) + list(ln.parameters()):
                p.grad = None
            xx = x.clone().requires_grad_(True)
            y = mod.rows_forward(xx, M_.Geom(B, H, H), ln, scale, True)
            (y.float() * R).sum().backward()
            torch.cuda.synchronize()
            outs.append((y.detach().float(), xx.grad.float(), {k: p.grad.clone() for k, p in list(mod.named_parameters()) +
                                                                [("ln." + k, p) for k, p in ln.named_parameters()]}))
        finally:
            OF.MLP_FUSED = True
    (y1, dx1, g1), (y2, dx2, g2) = outs
    rms = lambda t: float(t.pow(2).mean().sqrt())  # noqa: E731
    assert float((y1 - y2).abs().max()) <= 2e-2 * (float(y2.abs().max()))
    assert float((dx1 - dx2).abs().max()) <= 2e-2 * float(dx2.abs().max()) + 1e-6
    for k in g2:
        assert float((g1[k] - g2[k]).norm()) <= 2e-2 * float(g2[k].norm()) + 1e-6 * g2[k].numel() ** 0.5, k
