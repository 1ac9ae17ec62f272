"""The drop-in itself: `og.install()` rebinds the block classes inside an UNMODIFIED
reference checkout, after which the reference's own callers -- `scripts.train.build_model`, `MaxOutNet`,
`OutlookerFrontGridNet`, `src.training.train_full_model.train_model` -- and the reference's own tests run on
the sm_100a blocks.

CPU part (runs wherever a reference checkout is found: /root/reference in the authoring container,
baseline/_ref on the GPU box): structure, state_dict equality, strict load.
GPU part (-m gpu): the reference's tests/test_blocks.py and tests/test_models.py test FUNCTIONS, imported from
the reference and executed unmodified with CUDA as the default device, plus its training smoke test on cuda.
"""
import importlib
import importlib.util
import sys

import pytest
import torch

import outlook_grid_vision_transformer_b200 as og
from outlook_grid_vision_transformer_b200 import modules as ogm
from outlook_grid_vision_transformer_b200.config import CONFIG_DIR

REF = og.find_reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference checkout (neither /root/reference nor baseline/_ref)")


@pytest.fixture
def dropin():
    root = og.install()
    try:
        yield root
    finally:
        og.uninstall()


def _ref_build_model(model_cfg):
    train = importlib.import_module("scripts.train")
    return train.build_model(model_cfg)


@needs_ref
@pytest.mark.parametrize("yaml_name", ["cifar100_model_a_7m.yaml", "cifar100_model_a_14m.yaml", "tinyimagenet200_model_a.yaml",
                                       "cifar100_model_b.yaml"])
def test_install_rebinds_reference_models(yaml_name):
    """reference build_model -> reference MaxOutNet / OutlookerFrontGridNet made of OUR blocks; identical
    state_dict keys and shapes; a reference checkpoint loads strict=True; uninstall restores the reference."""
    mcfg = og.load_yaml(CONFIG_DIR / yaml_name)["model"]
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    sys.dont_write_bytecode = True
    torch.manual_seed(0)
    ref_model = _ref_build_model(mcfg)  # pure reference
    ref_blk = importlib.import_module("src.model.Out_Grid_Block").OutGridBlock
    assert ref_blk is not ogm.OutGridBlock
    og.install()
    try:
        assert importlib.import_module("src.Model_A_OutGridNet").OutGridBlock is ogm.OutGridBlock
        assert importlib.import_module("src.Model_B_OutGridNet").GridOnlyBlock is ogm.GridOnlyBlock
        model = _ref_build_model(mcfg)
        assert type(model).__module__.startswith("src.")  # the caller is the reference's own class
        blocks = [b for st in model.stages for b in st]
        assert blocks and all(isinstance(b, (ogm.OutGridBlock, ogm.GridOnlyBlock)) for b in blocks)
        if hasattr(model, "front"):
            assert all(isinstance(b, ogm.OutlookerBlock2d) for b in model.front)
        sd_ref, sd = ref_model.state_dict(), model.state_dict()
        assert list(sd) == list(sd_ref)
        assert all(sd[k].shape == sd_ref[k].shape and sd[k].dtype == sd_ref[k].dtype for k in sd)
        missing = model.load_state_dict(sd_ref, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        # weight-decay grouping of the reference loop keys on parameter names: must be unchanged
        warm = importlib.import_module("src.training.warmup")
        g_ref = warm.build_param_groups_no_wd(ref_model, 0.05)
        g_new = warm.build_param_groups_no_wd(model, 0.05)
        assert [len(g["params"]) for g in g_ref] == [len(g["params"]) for g in g_new]
    finally:
        og.uninstall()
    assert importlib.import_module("src.Model_A_OutGridNet").OutGridBlock is ref_blk
    assert not og.dropin.installed()


@needs_ref
def test_dropin_has_no_cpu_fallback(dropin):
    mcfg = og.load_yaml(CONFIG_DIR / "cifar100_model_a_7m.yaml")["model"]
    model = _ref_build_model(mcfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.randn(1, 3, 32, 32))


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture
def cuda_default():
    """CUDA as the default device, and fp32 products on the exact FFMA engine: the reference's own tolerance
    (atol 1e-6, rtol 1e-5, tests/test_blocks.py:71) is tighter than the 2^-17 of the three-bf16-plane tcgen05 scheme
    that fp32 mode uses by default (north_star bar for fp32: rtol 1e-3) -- `OGV_FP32_EXACT=1` is the library's switch."""
    from outlook_grid_vision_transformer_b200 import ops

    prev, prev_tc = torch.get_default_device(), ops.FP32_ON_TENSOR_CORES
    torch.set_default_device("cuda")
    ops.FP32_ON_TENSOR_CORES = False
    try:
        yield
    finally:
        torch.set_default_device(prev)
        ops.FP32_ON_TENSOR_CORES = prev_tc


def _ref_test_module(name):
    """Import tests/<name>.py of the REFERENCE under a private module name (this repo has its own `tests`)."""
    spec = importlib.util.spec_from_file_location(f"_ref_tests_{name}", REF / "tests" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("test_file,test_fn", [
    ("test_blocks", "test_grid_partition_roundtrip"),
    ("test_blocks", "test_outlook_attention_shapes"),
    ("test_blocks", "test_outgrid_block_forward_matches_manual"),
    ("test_models", "test_model_a_forward"),
    ("test_models", "test_model_b_forward"),
])
def test_reference_own_tests_pass_on_the_dropin(dropin, cuda_default, test_file, test_fn):
    """reference tests/test_blocks.py:32-71, tests/test_models.py:58-84, unmodified, default device = cuda."""
    import importlib.util  # noqa: F401

    mod = _ref_test_module(test_file)
    # the reference tests bound the class names at import: they must be OUR classes now
    if hasattr(mod, "OutGridBlock"):
        assert mod.OutGridBlock is ogm.OutGridBlock
    getattr(mod, test_fn)()


@needs_ref
@pytest.mark.gpu
def test_reference_train_loop_runs_on_the_dropin(dropin, tmp_path):
    """reference tests/test_training_smoke.py:32-75 with device='cuda' and bf16 autocast: the reference's
    train_model / train_one_epoch / evaluate_one_epoch drive the CUDA blocks end to end, checkpoints included."""
    from torch.utils.data import DataLoader, TensorDataset

    smoke = _ref_test_module("test_training_smoke")
    train_model = importlib.import_module("src.training.train_full_model").train_model
    MaxOutNet = importlib.import_module("src.Model_A_OutGridNet").MaxOutNet
    torch.manual_seed(0)
    model = MaxOutNet(num_classes=10, stages=[smoke._make_stage()], stem_dim=16, dpr_max=0.1)
    assert isinstance(model.stages[0][0], ogm.OutGridBlock)
    images, labels = torch.randn(8, 3, 8, 8), torch.randint(0, 10, (8,))
    tl = DataLoader(TensorDataset(images, labels), batch_size=4, shuffle=False)
    vl = DataLoader(TensorDataset(images, labels), batch_size=4, shuffle=False)
    save, last = tmp_path / "best.pt", tmp_path / "last.pt"
    history, trained = train_model(model=model, train_loader=tl, epochs=2, val_loader=vl, device="cuda", lr=1e-3,
                                   weight_decay=0.05, autocast_dtype="bf16", use_amp=True, grad_clip_norm=1.0,
                                   warmup_ratio=0.0, min_lr=0.0, label_smoothing=0.1, print_every=0,
                                   save_path=str(save), last_path=str(last), resume_path=None, mixup_alpha=0.0,
                                   cutmix_alpha=0.0, mix_prob=0.0, num_classes=10, channels_last=True, early_stop=False)
    assert len(history["train_loss"]) == 2 and all(map(lambda v: v == v, history["train_loss"]))
    assert save.exists() and last.exists()
    # the checkpoint the reference wrote loads strict=True into a PURE reference model (same keys, same shapes)
    og.uninstall()
    pure = importlib.import_module("src.Model_A_OutGridNet").MaxOutNet(num_classes=10, stages=[smoke._make_stage()],
                                                                     stem_dim=16, dpr_max=0.1)
    ck = torch.load(last, map_location="cpu", weights_only=False)
    pure.load_state_dict(ck["model"], strict=True)
