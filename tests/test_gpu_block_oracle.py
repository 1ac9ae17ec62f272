"""-m gpu: OutGridBlock / whole-model parity against the CPU oracle on the same seeded inputs, at
the stage shapes of the BASELINE configs (sizes the oracle finishes in seconds), plus
size-independent properties at the full BASELINE batch."""
import contextlib
from types import SimpleNamespace

import pytest
import torch

from oracle import outgrid_oracle as O
from oracle_cases import assert_close, assert_close_rms, assert_grad_close_bf16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2}

# (C, H, heads, outlook_heads, grid, B): stage shapes of cfg 1 (7M) and cfg 2/3 (14M @32 and @64)
STAGE_SHAPES = [
    (48, 32, 2, 2, 8, 2), (96, 16, 3, 3, 8, 3), (192, 8, 6, 6, 4, 4), (256, 4, 8, 8, 2, 6),
    (64, 32, 2, 2, 8, 2), (128, 16, 4, 4, 8, 3), (256, 8, 8, 8, 4, 4), (384, 4, 6, 6, 2, 6),
    (64, 64, 2, 2, 8, 1), (384, 8, 6, 6, 2, 2),
    (128, 32, 4, 4, 8, 1), (256, 16, 8, 8, 4, 2),  # the two remaining cfg 3 / cfg 4 stage shapes (64 px input)
]


def _block_and_oracle(C, H, heads, oheads, g, B, drop_path, seed):
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(seed)
    cfg = og.StageCfg(dim=C, depth=1, num_heads=heads, grid_size=g, outlook_heads=oheads, drop_path=drop_path)
    blk = og.OutGridBlock(cfg)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if "norm" in n or ".1." in n:  # LN / BN affine: move away from (1, 0)
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    x = torch.randn(B, C, H, H, generator=gen)
    R = torch.randn(B, C, H, H, generator=gen)
    return cfg, blk, x, R


def _oracle_block(cfg, blk, x, R, scales, dtype=torch.float64):
    params = {"m." + k: v.detach().to(dtype).clone().requires_grad_("running" not in k)
              for k, v in blk.state_dict().items() if v.is_floating_point()}
    params.update({"m." + k: v for k, v in blk.state_dict().items() if not v.is_floating_point()})
    xo = x.to(dtype).clone().requires_grad_(True)
    aux = {}
    y = O.outgrid_block(xo, params, "m", SimpleNamespace(**cfg.__dict__), True, aux, scales)
    (y * R.to(dtype)).sum().backward()
    grads = {k[2:]: p.grad for k, p in params.items() if p.is_floating_point() and p.requires_grad}
    return y.detach(), xo.grad, grads, {k[2:]: v for k, v in aux.items()}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("C,H,heads,oheads,g,B", STAGE_SHAPES)
def test_outgrid_block_train_matches_oracle(C, H, heads, oheads, g, B, dtype):
    from gpu_common import forced_drop_scales
    cfg, blk, x, R = _block_and_oracle(C, H, heads, oheads, g, B, drop_path=0.25, seed=C + H)
    gen = torch.Generator().manual_seed(5)
    scales = [((torch.rand(B, generator=gen) < 0.75).double() / 0.75) for _ in range(4)]
    yo, dxo, go, aux = _oracle_block(cfg, blk, x, R, scales)
    blk = blk.to(DEV).train()
    xg = x.to(DEV, dtype).requires_grad_(True)
    with forced_drop_scales(scales):
        y = blk(xg)
    assert y.shape == x.shape and y.dtype == dtype
    (y.float() * R.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    rtol = RTOL[dtype]
    assert_close(y.float(), yo, rtol, "forward")
    bf = dtype == torch.bfloat16
    assert_close_rms(y.float(), yo, rtol, "forward (elementwise rtol, atol = rtol * rms)",
                     outlier_frac=1e-4 if bf else 0.0, outlier_band=1.5 if bf else 1.0)
    assert_close(xg.grad.float(), dxo, rtol, "dx", atol=1e-6)
    for k, p in blk.named_parameters():
        assert p.grad is not None, f"no gradient for {k}"
        if dtype == torch.bfloat16:
            assert_grad_close_bf16(p.grad, go[k], rtol, f"grad[{k}]")
        else:
            assert_close(p.grad, go[k], rtol, f"grad[{k}]", atol=1e-5)
    for k, v in aux.items():
        assert_close(blk.state_dict()[k].double(), v.double(), rtol, f"buffer[{k}]")


@pytest.mark.parametrize("yaml_name,img,B", [("cifar100_model_a_7m.yaml", 32, 4), ("cifar100_model_a_14m.yaml", 32, 2)])
def test_model_fp32_logits_and_grads_match_oracle(yaml_name, img, B):
    """BASELINE config 1 (and the 14M net): whole-model logits + every gradient in fp32, rtol 1e-3."""
    import outlook_grid_vision_transformer_b200 as og
    from outlook_grid_vision_transformer_b200.config import CONFIG_DIR
    mcfg = dict(og.load_yaml(CONFIG_DIR / yaml_name)["model"], dpr_max=0.0)
    torch.manual_seed(7)
    model = og.build_model(mcfg)
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(B, 3, img, img, generator=gen)
    labels = torch.randint(0, mcfg["num_classes"], (B,), generator=gen)
    params = {k: (v.detach().double().clone().requires_grad_("running" not in k) if v.is_floating_point() else v)
              for k, v in model.state_dict().items()}
    logits_o = O.model_forward(x.double(), params, mcfg, True, {}, None)
    torch.nn.functional.cross_entropy(logits_o, labels).backward()
    model = model.to(DEV).train()
    logits = model(x.to(DEV).contiguous(memory_format=torch.channels_last))
    torch.nn.functional.cross_entropy(logits, labels.to(DEV)).backward()
    torch.cuda.synchronize()
    assert_close(logits, logits_o.detach(), 1e-3, "logits")
    for k, p in model.named_parameters():
        assert_close(p.grad, params[k].grad, 1e-3, f"grad[{k}]", atol=1e-6)


def test_full_batch_properties_bf16():
    """BASELINE config 2, stage-0 block at the full per-GPU batch (1024 x 64 x 32 x 32, bf16):
    (1) eval-mode outputs of individual samples equal the oracle run on those samples alone;
    (2) training-mode output is equivariant to a permutation of the batch (BN statistics are
    permutation invariant), i.e. sample b of blk(x[perm]) == sample perm[b] of blk(x)."""
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(0)
    B, C, H = 1024, 64, 32
    cfg = og.StageCfg(dim=C, depth=1, num_heads=2, grid_size=8, outlook_heads=2, drop_path=0.0)
    blk = og.OutGridBlock(cfg).to(DEV)
    x = torch.randn(B, C, H, H, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    blk.eval()
    with torch.no_grad():
        y = blk(x)
    params = {"m." + k: (v.detach().double().cpu() if v.is_floating_point() else v.cpu()) for k, v in blk.state_dict().items()}
    for b in (0, 511, 1023):
        yo = O.outgrid_block(x[b:b + 1].double().cpu(), params, "m", SimpleNamespace(**cfg.__dict__), False, None, None)
        assert_close(y[b:b + 1].float(), yo, 2e-2, f"eval sample {b}")
    blk.train()
    perm = torch.randperm(B, device=DEV)
    with torch.no_grad():
        y1 = blk(x)
        y2 = blk(x[perm].contiguous(memory_format=torch.channels_last))
    assert_close(y2.float(), y1[perm].float(), 2e-2, "batch-permutation equivariance")
