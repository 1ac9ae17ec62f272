#!/usr/bin/env python
"""Stage the UNMODIFIED reference where a GPU box can import it: baseline/_ref (git-ignored, not
gpurun-ignored, so it travels with the snapshot).  Only the importable Python tree is staged
(src/, scripts/, configs/, tests/ -- about 350 KB); notebooks, logs and result pickles stay behind.
Nothing under baseline/_ref is product code: bench.py's `eager_gpu` comparator and the drop-in tests
import it as the comparator / the caller, exactly as they would import /root/reference.

    python tools/install_reference.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DST = ROOT / "baseline" / "_ref"
PARTS = ("src", "scripts", "configs", "tests", "requirements.txt")


def install(src: Path = Path("/root/reference"), quiet: bool = False) -> bool:
    if not (src / "src" / "model" / "Out_Grid_Block.py").exists():
        if not quiet:
            print(f"reference not found at {src}; nothing staged", file=sys.stderr)
        return False
    DST.mkdir(parents=True, exist_ok=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb", "*.pkl", "*.pt", "*.png", "*.jpg")
    for part in PARTS:
        s, d = src / part, DST / part
        if s.is_dir():
            if d.exists():
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=ignore)
        elif s.exists():
            shutil.copy2(s, d)
    if not quiet:
        print(f"staged {src} -> {DST}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if install(Path(a.src)) else 1)
