"""-m gpu: SURVEY section 8 row f2 -- the conv -> BatchNorm -> act units around the blocks (ConvStem, stem_head.py:23-32;
Downsample, downsampling.py:28-65) with their BatchNorm + activation forward / backward on this package's kernels,
against the same nn.Sequential run by PyTorch (the reference's code path for these units): outputs, input / weight /
BatchNorm gradients, running statistics and the batch counter.  fp32 rtol 1e-3, bf16 (under autocast) rtol 2e-2."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle_cases import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _units():
    from outlook_grid_vision_transformer_b200.model import ConvStem, Downsample, DownsampleConfig
    return {
        "stem_3_64": (lambda: ConvStem(3, 64), (5, 3, 32, 32)),
        "stem_3_24_relu": (lambda: ConvStem(3, 24, act="relu"), (3, 3, 9, 7)),
        "down_conv_64_128": (lambda: Downsample(64, 128), (4, 64, 16, 16)),
        "down_conv_odd": (lambda: Downsample(16, 40, DownsampleConfig(kind="conv", act="gelu")), (3, 16, 7, 5)),
        "down_pool_32_64": (lambda: Downsample(32, 64, DownsampleConfig(kind="pool")), (4, 32, 8, 8)),
    }


def _seq(unit):
    return unit.stem if hasattr(unit, "stem") else unit.op


@pytest.mark.parametrize("name", list(_units()))
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_bn_act_matches_torch_sequential(name, mode):
    make, shape = _units()[name]
    torch.manual_seed(3)
    unit = make().to(DEV).train()
    with torch.no_grad():
        bn = _seq(unit)[-2]
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    ref = copy.deepcopy(_seq(unit))          # plain PyTorch modules, same parameters and buffers
    rtol = 1e-3 if mode == "fp32" else 2e-2
    g = torch.Generator().manual_seed(4)
    x = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
        y = unit(x1)
        yr = ref(x2)
    assert y.shape == yr.shape and y.dtype == yr.dtype
    dy = torch.randn(y.shape, generator=g).to(DEV).to(y.dtype)
    y.backward(dy)
    yr.backward(dy)
    torch.cuda.synchronize()
    assert_close(y.float(), yr.float(), rtol, "output")
    assert_close(x1.grad, x2.grad, rtol, "dx")
    for (k, p), (_, q) in zip(_seq(unit).named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, rtol, f"grad[{k}]", atol=1e-6)
    for (k, b), (_, q) in zip(_seq(unit).named_buffers(), ref.named_buffers()):
        if b.is_floating_point():
            assert_close(b, q, 1e-3 if mode == "fp32" else 5e-3, f"buffer[{k}]")
        else:
            assert int(b) == int(q) == 1, k


def test_eval_mode_uses_running_statistics():
    from outlook_grid_vision_transformer_b200.model import Downsample
    torch.manual_seed(5)
    unit = Downsample(32, 64).to(DEV)
    with torch.no_grad():
        unit.op[1].running_mean.uniform_(-0.3, 0.3)
        unit.op[1].running_var.uniform_(0.5, 2.0)
    unit.eval()
    ref = copy.deepcopy(unit.op)
    x = torch.randn(3, 32, 8, 8, device=DEV).contiguous(memory_format=torch.channels_last)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y, yr = unit(x1), ref(x2)
    assert_close(y, yr, 1e-3, "eval output")
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy)
    assert_close(x1.grad, x2.grad, 1e-3, "eval dx")
    for (k, p), (_, q) in zip(unit.op.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, 1e-3, f"eval grad[{k}]", atol=1e-6)
    assert int(unit.op[1].num_batches_tracked) == 0
    assert torch.equal(unit.op[1].running_mean, ref[1].running_mean)


def test_hooks_fall_back_to_the_module_chain():
    """A forward hook on the conv or the BatchNorm must fire with the tensors the reference would hand it."""
    from outlook_grid_vision_transformer_b200.model import ConvStem
    torch.manual_seed(6)
    unit = ConvStem(3, 32).to(DEV).train()
    seen = {}
    def hook(m, i, o):
        seen["bn_in"] = i[0].shape  # returns None: the output is left alone

    h = unit.stem[1].register_forward_hook(hook)
    x = torch.randn(2, 3, 8, 8, device=DEV)
    y = unit(x)
    h.remove()
    assert seen["bn_in"] == (2, 32, 8, 8) and y.shape == (2, 32, 8, 8)
    y2 = unit(x)  # fused again
    assert_close(y2, y, 1e-3, "fused vs module chain")


def test_momentum_none_is_a_cumulative_average():
    from outlook_grid_vision_transformer_b200.model import ConvStem
    torch.manual_seed(7)
    unit = ConvStem(3, 16).to(DEV).train()
    unit.stem[1].momentum = None
    ref = copy.deepcopy(unit.stem)
    for i in range(3):
        x = torch.randn(4, 3, 8, 8, device=DEV) + i
        unit(x)
        ref(x)
    assert_close(unit.stem[1].running_mean, ref[1].running_mean, 1e-3, "running_mean")
    assert_close(unit.stem[1].running_var, ref[1].running_var, 1e-3, "running_var")
    assert int(unit.stem[1].num_batches_tracked) == 3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 3, 5, 7), (3, 1, 4, 4), (2, 4, 9, 3), (1, 7, 6, 6)])
def test_im2col3x3_is_unfold_bit_exact(shape, dtype):
    """ogv_im2col3x3 against F.unfold on integer-valued inputs: bit-exact, pad columns zero."""
    from outlook_grid_vision_transformer_b200 import ops
    B, C, H, W = shape
    g = torch.Generator().manual_seed(8)
    x = torch.randint(-8, 9, shape, generator=g).to(dtype).to(DEV).contiguous(memory_format=torch.channels_last)
    kpad = (9 * C + 7) // 8 * 8
    cols = ops.im2col3x3(x, kpad)
    want = torch.nn.functional.unfold(x.float(), 3, padding=1)            # [B, C*9, H*W], row c*9 + tap
    want = want.view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)  # column tap*C + c
    assert torch.equal(cols[:, :9 * C].float(), want)
    assert not cols[:, 9 * C:].any()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stem_as_patches_gemm(mode):
    """The network input needs no gradient: the stem convolution runs as patches x tcgen05 GEMM (forward and weight
    gradient); same outputs / gradients / buffers as the PyTorch module chain."""
    from outlook_grid_vision_transformer_b200 import functional as OF
    from outlook_grid_vision_transformer_b200.model import ConvStem
    torch.manual_seed(9)
    unit = ConvStem(3, 64).to(DEV).train()
    ref = copy.deepcopy(unit.stem)
    rtol = 1e-3 if mode == "fp32" else 2e-2
    x = torch.randn(6, 3, 32, 32, device=DEV).contiguous(memory_format=torch.channels_last)
    seen = []
    orig = OF.ops.im2col3x3
    OF.ops.im2col3x3 = lambda *a: (seen.append(1), orig(*a))[1]
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            y, yr = unit(x), ref(x)
    finally:
        OF.ops.im2col3x3 = orig
    assert seen, "the stem did not take the patches x GEMM path"
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy)
    assert_close(y.float(), yr.float(), rtol, "output")
    for (k, p), (_, q) in zip(unit.stem.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, rtol, f"grad[{k}]", atol=1e-6)
    for (k, b), (_, q) in zip(unit.stem.named_buffers(), ref.named_buffers()):
        if b.is_floating_point():
            assert_close(b, q, 1e-3 if mode == "fp32" else 5e-3, f"buffer[{k}]")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (3, 16, 4, 4), (2, 64, 16, 16), (1, 24, 9, 6), (5, 128, 8, 8)])
def test_im2col_col2im_vec_bit_exact(shape, stride, dtype):
    """ogv_im2col3x3_vec against F.unfold and ogv_col2im3x3_vec against F.fold (its transpose) on integer-valued
    inputs: bit-exact, odd sizes and both strides (downsampling.py:41-47 is stride 2)."""
    from outlook_grid_vision_transformer_b200 import ops
    B, C, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randint(-8, 9, shape, generator=g).to(dtype).to(DEV).contiguous(memory_format=torch.channels_last)
    cols = ops.im2col3x3_vec(x, stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    want = torch.nn.functional.unfold(x.float(), 3, padding=1, stride=stride)       # [B, C*9, Ho*Wo], row c*9 + tap
    want = want.view(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)  # column tap*C + c
    assert cols.shape == want.shape
    assert torch.equal(cols.float(), want)
    d = torch.randint(-4, 5, (B * Ho * Wo, 9 * C), generator=g).to(dtype).to(DEV)
    dx = ops.col2im3x3_vec(d, B, H, W, C, stride)
    dn = d.float().view(B, Ho * Wo, 9, C).permute(0, 3, 2, 1).reshape(B, C * 9, Ho * Wo)
    want_dx = torch.nn.functional.fold(dn, (H, W), 3, padding=1, stride=stride)      # [B, C, H, W]
    assert torch.equal(dx.float().view(B, H, W, C).permute(0, 3, 1, 2), want_dx)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_downsample_convolution_runs_on_own_kernels(mode):
    """Downsample's 3x3 stride-2 convolution takes the patches x GEMM route in all three directions (no library
    convolution in the step) and agrees with the module chain at the stage shapes of cfg 2 (scaled-down batch)."""
    from outlook_grid_vision_transformer_b200 import functional as OF
    from outlook_grid_vision_transformer_b200.model import Downsample
    rtol = 1e-3 if mode == "fp32" else 2e-2
    for cin, hw in ((64, 32), (128, 16), (256, 8)):
        torch.manual_seed(12)
        unit = Downsample(cin, 2 * cin).to(DEV).train()
        ref = copy.deepcopy(unit.op)
        x = torch.randn(4, cin, hw, hw, device=DEV).contiguous(memory_format=torch.channels_last)
        x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        seen = []
        names = ("im2col3x3_vec", "col2im3x3_vec", "conv3x3_fwd", "conv3x3_wgrad")
        orig = {n: getattr(OF.ops, n) for n in names}
        for n in names:
            setattr(OF.ops, n, (lambda n: lambda *a, **k: (seen.append(n), orig[n](*a, **k))[1])(n))
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
                y, yr = unit(x1), ref(x2)
            dy = torch.randn_like(y)
            y.backward(dy)
            yr.backward(dy)
        finally:
            for n in names:
                setattr(OF.ops, n, orig[n])
        # bf16: implicit GEMM (no patch matrix); fp32: materialised patches through the three-plane fp32 GEMM
        want = ["conv3x3_fwd", "conv3x3_wgrad", "col2im3x3_vec"] if mode == "bf16" else ["im2col3x3_vec", "col2im3x3_vec"]
        assert seen == want, seen
        assert_close(y.float(), yr.float(), rtol, f"output C={cin}")
        assert_close(x1.grad, x2.grad, rtol, f"dx C={cin}")
        for (k, p), (_, q) in zip(unit.op.named_parameters(), ref.named_parameters()):
            assert_close(p.grad, q.grad, rtol, f"grad[{k}] C={cin}", atol=1e-6)


@pytest.mark.parametrize("shape,stride", [((4, 64, 32, 32), 2), ((3, 128, 16, 16), 2), ((4, 256, 8, 8), 2), ((3, 256, 8, 8), 2),
                                          ((2, 64, 16, 16), 1), ((5, 64, 8, 8), 1), ((2, 64, 64, 64), 2), ((9, 128, 4, 4), 1)])
def test_implicit_conv_equals_materialised_patches(shape, stride):
    """ogv_conv3x3_fwd / ogv_conv3x3_wgrad (5-D TMA boxes of x as the GEMM operand) against the same GEMMs on the
    materialised patch matrix, integer-valued operands: bit-exact (every sum is exact in fp32), including batches that
    do not fill the last 128-pixel tile and images smaller than a tile."""
    from outlook_grid_vision_transformer_b200 import ops
    B, C, H, W = shape
    Co = 2 * C if C < 256 else 384
    g = torch.Generator().manual_seed(13)
    x = torch.randint(-4, 5, shape, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    w2 = torch.randint(-2, 3, (Co, 9 * C), generator=g).to(torch.bfloat16).to(DEV)
    if not ops.conv3x3_supported(x, Co, stride):
        pytest.skip("geometry not served by the implicit path")
    cols = ops.im2col3x3_vec(x, stride)
    Mo = cols.shape[0]
    want = torch.empty((Mo, Co), device=DEV, dtype=torch.bfloat16)
    ops.gemm(cols, w2, want)
    got = ops.conv3x3_fwd(x, w2, stride)
    assert torch.equal(got, want)
    assert torch.equal(got.float(), (cols.float() @ w2.float().t()).to(torch.bfloat16).float())
    dy = torch.randint(-2, 3, (Mo, Co), generator=g).to(torch.bfloat16).to(DEV)
    dw_want = dy.float().t() @ cols.float()
    dw = torch.zeros((Co, 9 * C), device=DEV)
    ops.conv3x3_wgrad(x, dy, dw, stride)
    assert torch.equal(dw, dw_want)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("training", [True, False], ids=["train", "eval"])
@pytest.mark.parametrize("shape,classes", [((6, 384, 4, 4), 100), ((5, 64, 2, 2), 200), ((16, 128, 8, 8), 10)])
def test_head_matches_torch_modules(shape, classes, training, mode):
    """BatchNorm2d -> global average pool -> Linear (Model_A_OutGridNet.py:52-53,65-67) on own kernels against the
    PyTorch modules: logits, input / BatchNorm / classifier gradients, running statistics."""
    from outlook_grid_vision_transformer_b200.model import _Backbone
    from outlook_grid_vision_transformer_b200.config import StageCfg
    from outlook_grid_vision_transformer_b200 import functional as OF
    torch.manual_seed(14)
    B, C, H, W = shape

    class Net(_Backbone):
        def __init__(self):
            super().__init__()
            self.head_norm = nn.BatchNorm2d(C)
            self.classifier = nn.Linear(C, classes)

    net = Net().to(DEV).train(training)
    with torch.no_grad():
        net.head_norm.weight.uniform_(0.5, 1.5)
        net.head_norm.bias.uniform_(-0.5, 0.5)
        net.head_norm.running_mean.uniform_(-0.2, 0.2)
        net.head_norm.running_var.uniform_(0.5, 1.5)
    ref = copy.deepcopy(net)
    rtol = 1e-3 if mode == "fp32" else 2e-2
    x = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    seen = []
    orig = OF.head
    OF.head = lambda *a, **k: (seen.append(1), orig(*a, **k))[1]
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            y = net._head(x1)
            yr = ref.classifier(ref.head_norm(x2).mean(dim=(2, 3)))
    finally:
        OF.head = orig
    assert seen, "the head did not take the fused path"
    assert y.shape == yr.shape == (B, classes) and y.dtype == yr.dtype
    dy = torch.randn(y.shape, device=DEV).to(y.dtype)
    y.backward(dy)
    yr.backward(dy)
    assert_close(y.float(), yr.float(), rtol, "logits")
    assert_close(x1.grad, x2.grad, rtol, "dx")
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, rtol, f"grad[{k}]", atol=1e-6)
    for (k, b), (_, q) in zip(net.named_buffers(), ref.named_buffers()):
        if b.is_floating_point():
            assert_close(b, q, 1e-3 if mode == "fp32" else 5e-3, f"buffer[{k}]")
        else:
            assert int(b) == int(q) == (1 if training else 0), k
