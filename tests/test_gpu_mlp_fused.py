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
