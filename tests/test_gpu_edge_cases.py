"""-m gpu: edge cases of the drop-in path -- empty and single-image batches, ragged spatial sizes,
input memory formats, CUDA-graph replay, stochastic-depth draws, error behaviour on the device."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _block(C=32, heads=2, oheads=2, g=2, drop_path=0.0, seed=0):
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(seed)
    cfg = og.StageCfg(dim=C, depth=1, num_heads=heads, grid_size=g, outlook_heads=oheads, drop_path=drop_path)
    return og.OutGridBlock(cfg).to(DEV).train()


def test_empty_batch_is_a_no_op_at_every_entry_point():
    """B = 0 / M = 0: every C-ABI call returns OK without launching anything (no division by zero, no
    zero-sized grid)."""
    from outlook_grid_vision_transformer_b200 import ops
    bf = torch.bfloat16
    e = lambda *shape, dt=bf: torch.empty(shape, device=DEV, dtype=dt)  # noqa: E731
    f = lambda *shape: torch.zeros(shape, device=DEV)  # noqa: E731
    C, Cm = 16, 64
    ops.gemm(e(0, C), e(Cm, C), e(0, Cm))
    ops.layernorm_fwd(e(0, C), f(C), f(C), 1e-5)
    ops.outlook_core_fwd(e(0, 40), 0, 8, 8, C, 2)
    ops.outlook_core_bwd(e(0, 40), e(0, C), 0, 8, 8, C, 2)
    ops.grid_attn_fwd(e(0, 3 * C), 0, 8, 8, C, 2, 2)
    ops.grid_attn_bwd(e(0, 3 * C), e(0, C), 0, 8, 8, C, 2, 2)
    ops.dwconv_fwd(e(0, Cm), f(Cm), f(Cm), f(Cm, 9), f(Cm), f(Cm), 0, 8, 8, "silu")
    ops.dwconv_bwd(e(0, Cm), e(0, Cm), f(Cm), f(Cm), f(Cm), f(Cm), f(Cm, 9), f(Cm, 9), f(Cm), f(Cm), 0, 8, 8, "silu")
    ops.se_pool(e(0, Cm), f(Cm), f(Cm), 0, 64, "silu")
    ops.bn_act_gate(e(0, Cm), f(Cm), f(Cm), f(0, Cm), 0, 64, "silu")
    ops.colsum(e(0, C), f(C))
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_single_image_batch_matches_first_image_of_eval_batch(dtype):
    """Eval mode (running statistics) has no cross-image coupling: image 0 alone == image 0 of a batch of 3."""
    blk = _block().eval()
    x = torch.randn(3, 32, 8, 8, device=DEV, dtype=dtype)
    with torch.no_grad():
        y3 = blk(x)
        y1 = blk(x[:1].contiguous())
    torch.testing.assert_close(y1.float(), y3[:1].float(), rtol=1e-5 if dtype == torch.float32 else 2e-2,
                               atol=1e-5 if dtype == torch.float32 else 2e-2)


def test_channels_last_and_contiguous_inputs_give_identical_results():
    blk = _block()
    x = torch.randn(2, 32, 8, 8, device=DEV)
    blk.eval()
    with torch.no_grad():
        a = blk(x)
        b = blk(x.contiguous(memory_format=torch.channels_last))
    assert torch.equal(a, b)
    assert a.shape == x.shape


def test_ragged_spatial_size_runs_and_matches_oracle_forward():
    """H != W, W not a multiple of the tile widths; grid size divides both."""
    from types import SimpleNamespace
    from oracle import outgrid_oracle as O
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(3)
    cfg = og.StageCfg(dim=16, depth=1, num_heads=2, grid_size=3, outlook_heads=2, drop_path=0.0)
    blk = og.OutGridBlock(cfg)
    x = torch.randn(2, 16, 6, 9)
    params = {"m." + k: (v.detach().double() if v.is_floating_point() else v) for k, v in blk.state_dict().items()}
    yo = O.outgrid_block(x.double(), params, "m", SimpleNamespace(**cfg.__dict__), True, {}, None)
    y = blk.to(DEV).train()(x.to(DEV))
    err = float((y.detach().double().cpu() - yo).abs().max() / yo.abs().max())
    assert err < 1e-3, f"ragged 6x9 forward rel err {err:.2e}"


def test_grid_size_must_divide_the_image():
    blk = _block(g=3)
    with pytest.raises(ValueError):
        blk(torch.randn(1, 32, 8, 8, device=DEV))


def test_cpu_tensor_is_rejected_loudly():
    blk = _block()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.randn(1, 32, 8, 8))


def test_drop_path_draws_follow_the_reference_order():
    """Stochastic depth consumes the global CUDA generator exactly like the reference's DropPath: one
    bernoulli_ per DropPath in the order outlook.dp1, outlook.dp2, dp2, dp3."""
    from outlook_grid_vision_transformer_b200 import modules as M
    blk = _block(drop_path=0.5)
    x = torch.randn(6, 32, 8, 8, device=DEV)
    seen = []
    orig = M._drop_scale

    def spy(dp, xx, batch):
        s = orig(dp, xx, batch)
        if s is not None:
            seen.append(s.clone())
        return s

    M._drop_scale = spy
    try:
        torch.manual_seed(123)
        blk(x)
    finally:
        M._drop_scale = orig
    torch.manual_seed(123)
    want = [(torch.empty((6, 1, 1, 1), device=DEV).bernoulli_(0.5) / 0.5).reshape(6) for _ in range(4)]
    assert len(seen) == 4
    for a, b in zip(seen, want):
        assert torch.equal(a, b)


def test_cast_batch_matches_per_weight_casts():
    """ogv_cast_batch: many (src -> dst, dst_t) casts in one launch == ogv_cast_transpose one by one, including
    ragged shapes, strided destination views, a missing transposed output and an fp32 copy item."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(5)
    shapes = [(64, 256), (40, 24), (1, 72), (129, 33), (256, 64)]
    triples, want = [], []
    for i, (r, c) in enumerate(shapes):
        src = torch.randn(r, c, device=DEV)
        big = torch.zeros(r + 3, c + 8, device=DEV, dtype=torch.bfloat16)
        big_t = torch.zeros(c, r + 8, device=DEV, dtype=torch.bfloat16)
        dst, dst_t = big[1:r + 1, :c], (big_t[:, :r] if i != 2 else None)
        triples.append((src, dst, dst_t))
        a, b = torch.zeros_like(big), torch.zeros_like(big_t)
        ops.cast_transpose(src, a[1:r + 1, :c], b[:, :r] if i != 2 else None)
        want.append((big, a, big_t, b))
    bias = torch.randn(1, 72, device=DEV)
    cat = torch.zeros(80, device=DEV)
    triples.append((bias, cat[:72].reshape(1, 72), None))
    ops.CastBatch(triples).run()
    for big, a, big_t, b in want:
        assert torch.equal(big, a) and torch.equal(big_t, b)
    assert torch.equal(cat[:72], bias[0]) and float(cat[72:].abs().sum()) == 0.0


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_standalone_squeeze_excite_forward_and_backward(dt):
    """SqueezeExcite used on its own (mbc_conv.py:9-27): forward AND every gradient against a PyTorch restatement."""
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(3)
    B, C, H = 3, 32, 8
    se = og.SqueezeExcite(C, se_ratio=0.25, act="silu").to(DEV)
    x = torch.randn(B, C, H, H, device=DEV).to(dt).requires_grad_(True)
    R = torch.randn(B, C, H, H, device=DEV)
    y = se(x)
    (y.float() * R).sum().backward()
    xr = x.detach().float().requires_grad_(True)
    w1 = se.fc1.weight.detach().clone().requires_grad_(True)
    b1 = se.fc1.bias.detach().clone().requires_grad_(True)
    w2 = se.fc2.weight.detach().clone().requires_grad_(True)
    b2 = se.fc2.bias.detach().clone().requires_grad_(True)
    s = torch.nn.functional.silu(torch.nn.functional.conv2d(xr.mean(dim=(2, 3), keepdim=True), w1, b1))
    yr = xr * torch.sigmoid(torch.nn.functional.conv2d(s, w2, b2))
    (yr * R).sum().backward()
    tol = 1e-3 if dt == torch.float32 else 2e-2
    close = lambda a, b, what: torch.testing.assert_close(a.float(), b, rtol=tol, atol=tol * float(b.abs().max()), msg=lambda m: f"{what}: {m}")  # noqa: E731
    close(y, yr.detach(), "forward")
    close(x.grad, xr.grad, "dx")
    close(se.fc1.weight.grad, w1.grad, "dW1")
    close(se.fc1.bias.grad, b1.grad, "db1")
    close(se.fc2.weight.grad, w2.grad, "dW2")
    close(se.fc2.bias.grad, b2.grad, "db2")


def test_forward_hook_on_outlook_logits_conv_fires_and_matches():
    """The attention-map tools hook `OutlookAttention2d.attn` (heat_map_att_outlooker.py:25-42): the hook must fire with
    the logits [B, heads*9, H, W] the reference conv would produce, and the block output must not change."""
    blk = _block(C=32, oheads=2).eval()
    x = torch.randn(2, 32, 8, 8, device=DEV)
    with torch.no_grad():
        want = blk(x)
    seen = {}
    h = blk.outlook.attn.attn.register_forward_hook(lambda m, inp, out: seen.update(inp=inp[0].detach(), out=out.detach()))
    try:
        with torch.no_grad():
            got = blk(x)
    finally:
        h.remove()
    assert "out" in seen and tuple(seen["out"].shape) == (2, 2 * 9, 8, 8)
    ref_logits = torch.nn.functional.conv2d(seen["inp"], blk.outlook.attn.attn.weight, blk.outlook.attn.attn.bias)
    torch.testing.assert_close(seen["out"], ref_logits, rtol=1e-4, atol=1e-5)
    xn = torch.nn.functional.layer_norm(x.permute(0, 2, 3, 1), (32,), blk.outlook.norm1.ln.weight, blk.outlook.norm1.ln.bias, 1e-6)
    torch.testing.assert_close(seen["inp"], xn.permute(0, 3, 1, 2), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4)
    with torch.no_grad():
        again = blk(x)  # hook removed: back on the fused path
    torch.testing.assert_close(again, want, rtol=0, atol=0)
