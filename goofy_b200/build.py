"""Build recipe for libgoofy_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libgoofy_b200.so"
SOURCES = [CSRC / "capi.cu"]
HEADERS = [CSRC / "lanes.cuh", CSRC / "block_codec.cuh", CSRC / "encode_kernels.cuh", CSRC / "tma_kernels.cuh", CSRC / "decode_kernels.cuh",
           CSRC / "block_decode.cuh", CSRC / "copy_pool.h", CSRC / "rgb_pack.h", CSRC / "hybrid_choice.h", CSRC / "host_neighbours.cuh",
           CSRC / "host_common.cuh", CSRC / "host_launch.cuh", CSRC / "host_resources.cuh", CSRC / "host_pipeline.cuh", CSRC / "host_batch.cuh",
           PKG_DIR.parent / "include" / "goofy_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "shared",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libgoofy_b200.so cannot be built (there is no CPU fallback)")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS if p.exists())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, "-Xptxas", "-v", "-o", str(LIB_PATH), *map(str, SOURCES)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build_library(force=True, verbose=True)
    print(LIB_PATH)
