"""Drop-in of the sm_100a blocks into an UNMODIFIED reference checkout.

The reference resolves its block classes through star-import chains (src/Model_A_OutGridNet.py:1-4,
src/Model_B_OutGridNet.py:6-9, src/model/Out_Grid_Block.py:5-7, src/model/Grid_Only_Block.py:5-9), so every
module of the chain holds its own binding of each class name.  `install()` rebinds those names, in every
module that has them, to the classes of this package; `MaxOutNet` / `OutlookerFrontGridNet` /
`scripts.train.build_model` then build their stages out of the CUDA blocks without a line of the reference
being edited.  `uninstall()` restores the original bindings.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path
from typing import Dict, Optional, Tuple

# classes of the hot path; the callers (MaxOutNet, ConvStem, Downsample, ...) stay the reference's
NAMES = ("OutGridBlock", "GridOnlyBlock", "OutlookerBlock2d", "OutlookAttention2d", "MBConv", "MBConvConfig",
         "SqueezeExcite", "GridAttention2D", "GridAttention2DConfig", "AttentionConfig", "MultiHeadSelfAttention",
         "MLP", "MLP2d", "LayerNorm2d", "DropPath")
# every module of the reference's import chain that binds (some of) those names
MODULES = ("src.model.outlook_attention", "src.model.Outlook_Block", "src.model.mbc_conv", "src.model.grid_attention",
           "src.model.Out_Grid_Block", "src.model.Grid_Only_Block", "src.Model_A_OutGridNet", "src.Model_B_OutGridNet")

_REPO = Path(__file__).resolve().parent.parent
_saved: Dict[Tuple[str, str], object] = {}


def find_reference_root(reference_root: Optional[os.PathLike] = None) -> Optional[Path]:
    """The reference checkout to bind into: the argument, $OGV_REFERENCE_ROOT, /root/reference, or the
    git-ignored copy tools/install_reference.py leaves under baseline/_ref (what travels to a GPU box)."""
    cands = [reference_root, os.environ.get("OGV_REFERENCE_ROOT"), "/root/reference", _REPO / "baseline" / "_ref"]
    for c in cands:
        if c and (Path(c) / "src" / "model" / "Out_Grid_Block.py").exists():
            return Path(c)
    return None


def install(reference_root: Optional[os.PathLike] = None) -> Path:
    """Rebind the reference's block classes to this package's.  Returns the reference root used."""
    import outlook_grid_vision_transformer_b200 as og

    root = find_reference_root(reference_root)
    if root is None:
        raise FileNotFoundError("no reference checkout found (argument, $OGV_REFERENCE_ROOT, /root/reference, baseline/_ref)")
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    sys.dont_write_bytecode = True  # the reference tree may be read-only
    for mname in MODULES:
        mod = importlib.import_module(mname)
        for name in NAMES:
            if hasattr(mod, name):
                _saved.setdefault((mname, name), getattr(mod, name))
                setattr(mod, name, getattr(og, name))
    return root


def uninstall() -> None:
    for (mname, name), obj in _saved.items():
        mod = sys.modules.get(mname)
        if mod is not None:
            setattr(mod, name, obj)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)
