"""Build libtio_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python torchio_b200/csrc/build.py [--force] [--verbose]

The shared library exports exactly the C-ABI of include/tio_b200.h and is
loaded with ctypes by torchio_b200/_native.py.  nvcc cross-compiles without a
GPU; the .so travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SOURCES = ["error.cu", "resample.cu", "resample_tile.cu", "intensity.cu", "fused_intensity.cu",
           "mt19937_jump.cpp", "mt19937.cu", "patches.cu"]
HEADERS = [HERE / "common.cuh", HERE / "intensity_common.cuh", HERE / "resample_common.cuh", ROOT / "include" / "tio_b200.h"]
OUT = HERE / "libtio_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    newest = max(p.stat().st_mtime for p in [*(HERE / s for s in SOURCES), *HEADERS])
    return OUT.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(OUT), *[str(HERE / s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed with exit code {proc.returncode}")
    (HERE / "ptxas.log").write_text(proc.stdout + proc.stderr)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
