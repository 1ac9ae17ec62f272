"""Stage configuration record and YAML loading (mirror of src/stage_config.py:4-34 and
scripts/train.py:23-30).  Unknown YAML keys raise TypeError exactly like `StageCfg(**yaml_stage)`."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List

import yaml

REPO_ROOT = Path(__file__).resolve().parent.parent
CONFIG_DIR = REPO_ROOT / "configs"


@dataclass
class StageCfg:
    # core dims
    dim: int
    depth: int
    # grid attention
    num_heads: int
    grid_size: int
    window_size: int = 8
    # outlooker
    outlook_heads: int = 6
    outlook_kernel: int = 3
    outlook_mlp_ratio: float = 2.0
    # MBConv
    mbconv_expand_ratio: float = 4.0
    mbconv_se_ratio: float = 0.25
    mbconv_act: str = "silu"
    use_bn: bool = True
    # drops
    attn_drop: float = 0.0
    proj_drop: float = 0.0
    ffn_drop: float = 0.0
    drop_path: float = 0.0
    # MLP (BHWC)
    mlp_ratio: float = 4.0
    mlp_act: str = "gelu"


def load_yaml(path) -> dict:
    with Path(path).open("r", encoding="utf-8") as f:
        return yaml.safe_load(f) or {}


def build_stages(stage_cfgs: List[dict]) -> List[StageCfg]:
    return [StageCfg(**cfg) for cfg in stage_cfgs]


# The five workloads BASELINE.json names (config file, image size, per-GPU batch, dtype, mode).
BASELINE_CONFIGS = {
    "cfg1_7m_32_fp32": dict(yaml="cifar100_model_a_7m.yaml", img=32, batch=128, dtype="fp32", mode="train"),
    "cfg2_14m_32_bf16": dict(yaml="cifar100_model_a_14m.yaml", img=32, batch=1024, dtype="bf16", mode="train"),
    "cfg3_14m_64_bf16": dict(yaml="cifar100_64_model_a.yaml", img=64, batch=256, dtype="bf16", mode="train"),
    "cfg4_22m_tin_64_bf16": dict(yaml="tinyimagenet200_model_a.yaml", img=64, batch=256, dtype="bf16", mode="train"),
    "cfg5_model_b_eval": dict(yaml="cifar100_model_b.yaml", img=32, batch=1024, dtype="bf16", mode="eval"),
    # not a BASELINE config: the 7M net measured the way the 14M headline is (bf16 autocast, batch 1024 per GPU)
    "cfg1b_7m_32_bf16": dict(yaml="cifar100_model_a_7m.yaml", img=32, batch=1024, dtype="bf16", mode="train"),
}
