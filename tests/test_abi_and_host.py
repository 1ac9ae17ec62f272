"""CPU suite (-m "not gpu"): the C-ABI library loads and exports every symbol include/ogv.h declares,
the ctypes table matches the header, and the host-side mirror of the reference interface behaves
like the reference (constructors, state_dict keys, error behaviour).  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import outlook_grid_vision_transformer_b200 as og
from outlook_grid_vision_transformer_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "ogv.h").read_text()


def _declared():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(ogv_[a-z0-9_]+)\s*\(", body)))


def test_header_declares_what_ctypes_binds():
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    for name in _declared():
        assert hasattr(lib, name), f"libogvit.so does not export {name}"
    assert lib.ogv_version() >= 100


def test_header_arity_matches_ctypes_table():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", body, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes table has {len(argtypes)}"


def test_gemm_args_struct_layout_is_c_compatible():
    # every pointer/long long 8 bytes, ints 4: the struct must not exceed the C layout computed by hand
    assert ctypes.sizeof(_lib.GemmArgs) % 8 == 0
    assert _lib.GemmArgs.A.offset == 0 and _lib.GemmArgs.M.offset == 64


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    rc = lib.ogv_layernorm_fwd(None, None, None, None, None, None, 4, 16, 1e-5, 0, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = lib.ogv_grid_attn_fwd(ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 6, 8, 16, 4, 4, 0, None)
    assert rc == -1 and "divisible by grid_size" in _lib.last_error()


def test_no_cpu_fallback():
    blk = og.OutGridBlock(og.StageCfg(dim=16, depth=1, num_heads=4, grid_size=2, outlook_heads=4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.randn(2, 16, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk.mlp(torch.randn(2, 8, 8, 16))


def test_constructor_validation_mirrors_reference():
    with pytest.raises(AssertionError):
        og.OutlookAttention2d(10, num_heads=3)
    with pytest.raises(ValueError):
        og.OutlookAttention2d(16, num_heads=4, kernel_size=4)
    with pytest.raises(ValueError):
        og.OutlookAttention2d(16, num_heads=4, stride=0)
    with pytest.raises(ValueError):
        og.SqueezeExcite(16, se_ratio=0.0)
    with pytest.raises(ValueError):
        og.MBConv(0, 16)
    with pytest.raises(ValueError):
        og.MBConv(16, 16, stride=3)
    with pytest.raises(ValueError):
        og.MultiHeadSelfAttention(og.AttentionConfig(dim=10, num_heads=3))
    with pytest.raises(ValueError):
        og.GridAttention2D(og.GridAttention2DConfig(mode="window", dim=16, num_heads=4, grid_size=2))
    with pytest.raises(ValueError):
        og.make_activation("tanh")
    with pytest.raises(TypeError):
        og.StageCfg(dim=16, depth=1, num_heads=4, grid_size=2, not_a_field=1)
    with pytest.raises(ValueError):
        og.grid_partition(torch.zeros(1, 6, 8, 4), 4)


def test_grid_partition_roundtrip_exact():
    x = torch.randn(2, 8, 8, 6)
    grids, meta = og.grid_partition(x, grid_size=2)
    assert grids.shape == (8, 4, 4, 6) and meta == (2, 8, 8, 6, 2)
    assert torch.equal(og.grid_unpartition(grids, meta), x)


def test_block_attribute_surface_and_state_dict_keys():
    blk = og.OutGridBlock(og.StageCfg(dim=16, depth=1, num_heads=4, grid_size=2, outlook_heads=4, drop_path=0.1))
    for attr in ("outlook", "mbconv", "norm2", "grid_attn", "dp2", "norm3", "mlp", "dp3"):
        assert hasattr(blk, attr)
    keys = set(blk.state_dict())
    for k in ("outlook.norm1.ln.weight", "outlook.attn.attn.weight", "outlook.attn.v.bias", "outlook.attn.proj.weight",
              "outlook.mlp.fc1.weight", "mbconv.expand.0.weight", "mbconv.expand.1.running_var",
              "mbconv.depthwise.0.weight", "mbconv.depthwise.1.num_batches_tracked", "mbconv.se.fc1.weight",
              "mbconv.project.1.weight", "norm2.weight", "grid_attn.mhsa.qkv.weight", "grid_attn.mhsa.proj.bias",
              "norm3.bias", "mlp.fc1.weight", "mlp.fc2.bias"):
        assert k in keys, k
    sd = blk.state_dict()
    assert sd["outlook.attn.attn.weight"].shape == (36, 16, 1, 1)
    assert sd["mbconv.depthwise.0.weight"].shape == (64, 1, 3, 3)
    assert sd["grid_attn.mhsa.qkv.weight"].shape == (48, 16)
    assert type(blk.grid_attn).__name__ == "GridAttention2D" and type(blk.grid_attn.mhsa).__name__ == "MultiHeadSelfAttention"
    assert type(blk.outlook.attn).__name__ == "OutlookAttention2d"


@pytest.mark.parametrize("name,nparams", [("cifar100_model_a_7m.yaml", 7518102), ("cifar100_model_a_14m.yaml", 14599198),
                                          ("cifar100_64_model_a.yaml", 14599198),
                                          ("tinyimagenet200_model_a.yaml", 22542628), ("cifar100_model_b.yaml", 12266266)])
def test_baseline_configs_build_with_reference_param_counts(name, nparams):
    from outlook_grid_vision_transformer_b200.config import CONFIG_DIR
    model = og.build_model(og.load_yaml(CONFIG_DIR / name)["model"])
    assert sum(p.numel() for p in model.parameters()) == nparams


def test_drop_path_schedule_and_identity_for_first_block():
    from outlook_grid_vision_transformer_b200.config import CONFIG_DIR
    model = og.build_model(og.load_yaml(CONFIG_DIR / "cifar100_model_a_14m.yaml")["model"])
    blocks = [b for st in model.stages for b in st]
    assert isinstance(blocks[0].dp2, torch.nn.Identity) and isinstance(blocks[0].outlook.dp1, torch.nn.Identity)
    assert abs(blocks[-1].dp3.drop_prob - 0.08) < 1e-12
    assert abs(blocks[1].dp2.drop_prob - 0.08 / 7) < 1e-12


def test_state_dict_matches_golden_reference_keys():
    from oracle_cases import load_golden
    case = load_golden()["model_a_tiny"]
    model = og.build_model(case["model_cfg"])
    assert list(model.state_dict().keys()) == list(case["state"].keys())
    model.load_state_dict(case["state"], strict=True)


def test_weight_copies_are_rebuilt_every_training_forward():
    """torch's fused AdamW updates parameters without bumping their version counter, so in training mode the
    compute-dtype weight copies must never come from a cache (regression: stale bf16 weights after opt.step())."""
    from outlook_grid_vision_transformer_b200.modules import _Prep
    w = torch.nn.Parameter(torch.zeros(2, 2))
    calls = []
    prep = _Prep()
    build = lambda: calls.append(1) or len(calls)  # noqa: E731
    assert prep.get(True, [w], torch.bfloat16, build) == 1
    assert prep.get(True, [w], torch.bfloat16, build) == 2      # training: always fresh
    assert prep.get(False, [w], torch.bfloat16, build) == 3     # eval: first call after the mode switch rebuilds ...
    assert prep.get(False, [w], torch.bfloat16, build) == 3     # ... then hits the cache
    with torch.no_grad():
        w.add_(1.0)                                             # versioned in-place update invalidates it
    assert prep.get(False, [w], torch.bfloat16, build) == 4


def test_cast_item_record_matches_the_c_struct(tmp_path):
    """ops.CastBatch writes ogv_cast_item records with numpy: the field offsets must be the C compiler's."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ogv.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ogv_cast_item), offsetof(ogv_cast_item, src),'
                   'offsetof(ogv_cast_item, dst), offsetof(ogv_cast_item, dst_t), offsetof(ogv_cast_item, ld_dst),'
                   'offsetof(ogv_cast_item, ld_dst_t), offsetof(ogv_cast_item, rows), offsetof(ogv_cast_item, cols),'
                   'offsetof(ogv_cast_item, tile0), offsetof(ogv_cast_item, dst_dtype));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", str(root / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [56, 0, 8, 16, 24, 32, 40, 44, 48, 52]


def test_bulk_refreshed_weight_copies_are_handed_out_only_while_fresh():
    """_Prep in training mode: per-layer rebuild by default; the recorded buffers are reused only between
    refresh_prepared() and the end of the forward (`_BULK_FRESH`), and never for other parameters / dtypes."""
    from outlook_grid_vision_transformer_b200 import functional as OF
    from outlook_grid_vision_transformer_b200 import modules as M
    w = torch.nn.Parameter(torch.zeros(2, 2))
    w2 = torch.nn.Parameter(torch.zeros(2, 2))
    calls = []

    def build():
        calls.append(1)
        if OF._CAST_LOG is not None:  # training-mode builds record their casts for refresh_prepared
            OF._CAST_LOG.append(("src", "dst", None))
        return len(calls)

    prep = M._Prep()
    assert prep.get(True, [w], torch.bfloat16, build) == 1
    assert prep._bulk is not None and prep._bulk[2] == [("src", "dst", None)]
    assert OF._CAST_LOG is None                                   # the recorder is off outside a build
    assert prep.get(True, [w], torch.bfloat16, build) == 2        # not fresh: rebuilt
    M._BULK_FRESH = True
    try:
        assert prep.get(True, [w], torch.bfloat16, build) == 2    # fresh: the recorded value, no rebuild
        assert prep.get(True, [w2], torch.bfloat16, build) == 3   # other parameter storage: rebuilt
        assert prep.get(True, [w2], torch.float16, build) == 4    # other dtype: rebuilt
        assert prep.get(False, [w2], torch.float16, build) == 5   # eval mode has its own cache
        assert prep.get(True, [w2], torch.float16, build) == 4    # ... which does not disturb the bulk record
    finally:
        M._BULK_FRESH = False
    assert prep.get(True, [w2], torch.float16, build) == 6


def test_implicit_convolution_geometry_gate():
    """ogv_conv3x3_supported is host logic (no GPU): the implicit-GEMM convolution serves exactly the geometries whose
    128- and 64-pixel tiles are whole output rows / whole images of a channels_last image with Cin % 64 == 0 --
    every Downsample unit of the BASELINE configs -- and sends everything else to the materialised-patch route."""
    from outlook_grid_vision_transformer_b200 import _lib
    ok = _lib.lib().ogv_conv3x3_supported
    # (B, H, W, Cin, Co, stride): the Downsample units of cfg 2 (32 px) and cfg 3 / 4 (64 px), batch 1024 / 256 / 1
    for geom in [(1024, 32, 32, 64, 128, 2), (1024, 16, 16, 128, 256, 2), (1024, 8, 8, 256, 384, 2),
                 (256, 64, 64, 64, 128, 2), (256, 32, 32, 128, 256, 2), (1, 16, 16, 256, 512, 2), (3, 8, 8, 64, 64, 1)]:
        assert ok(*geom) == 1, geom
    for geom, why in [((8, 32, 32, 48, 96, 2), "Cin % 64 (the 7M net: materialised patches)"),
                      ((8, 31, 32, 64, 128, 2), "odd height at stride 2"),
                      ((8, 48, 48, 64, 128, 2), "24 output columns do not tile 128 pixels"),
                      ((8, 32, 32, 64, 100, 2), "Co % 8"),
                      ((8, 32, 32, 64, 128, 3), "stride"),
                      ((8, 12, 12, 64, 128, 2), "36-pixel images do not tile 128 pixels")]:
        assert ok(*geom) == 0, (geom, why)


def test_cross_entropy_falls_through_to_torch_off_the_gpu_path():
    """og.cross_entropy is the ogv_xent kernels for [B, K] CUDA logits with int64 class indices; every other signature
    F.cross_entropy accepts goes to PyTorch unchanged (here: CPU tensors, probabilities as targets)."""
    import torch.nn.functional as F
    import outlook_grid_vision_transformer_b200 as og
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 10, generator=g, requires_grad=True)
    y = torch.randint(0, 10, (6,), generator=g)
    a = og.cross_entropy(x, y, label_smoothing=0.1)
    b = F.cross_entropy(x, y, label_smoothing=0.1)
    assert torch.equal(a, b)
    a.backward()
    p = torch.softmax(torch.randn(6, 10, generator=g), dim=1)
    assert torch.equal(og.cross_entropy(x, p), F.cross_entropy(x, p))
