"""Build libogvit.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m outlook_grid_vision_transformer_b200.build [--force] [--verbose]

The library has no link-time dependency on libcuda / torch: the TMA descriptor encoder is
resolved at run time through cudaGetDriverEntryPoint.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libogvit.so"

SOURCES = [
    "ogv_misc.cu",
    "ogv_gemm_simt.cu",
    "ogv_gemm_tc.cu",
    "ogv_norm.cu",
    "ogv_outlook.cu",
    "ogv_mbconv.cu",
    "ogv_dwconv.cu",
    "ogv_tma.cu",
    "ogv_gridattn.cu",
    "ogv_optim.cu",
    "ogv_mlp_fused.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; set NVCC or add /usr/local/cuda/bin to PATH")
    return cand


def _newest_input_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "ogv.h", Path(__file__)]
    return max(f.stat().st_mtime for f in files if f.exists())


def needs_build() -> bool:
    return (not LIB_PATH.exists()) or LIB_PATH.stat().st_mtime < _newest_input_mtime()


def _compile_one(src: str, verbose: bool) -> Path:
    obj = BUILD_DIR / (Path(src).stem + ".o")
    extra = os.environ.get("OGV_NVCC_EXTRA", "").split()  # e.g. -DOGV_EPI_WARPS=16 for A/B measurements
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}\n")
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile + link in-tree.  Guarded by an exclusive file lock and finished with an atomic rename, so the ranks of a
    multi-process launch that all find the library missing cannot interleave writes to _build/*.o or load a
    half-written libogvit.so: the first one builds, the others wait and find it up to date."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl

    BUILD_DIR.mkdir(exist_ok=True)
    with open(BUILD_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> Path:
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(lambda s: _compile_one(s, verbose), SOURCES))
    tmp = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [_nvcc(), "-shared", "-o", str(tmp), *map(str, objs),
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}\n")
        raise RuntimeError("link of libogvit.so failed")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
