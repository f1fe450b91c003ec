"""Build libg4s_rasterizer.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

No libtorch, no pybind: the library only depends on the CUDA runtime, so it compiles in
seconds and is ABI-stable across torch versions.  `python -m g4splat_b200.build [--force]`.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "libg4s_rasterizer.so"
SOURCES = ["api.cu", "project.cu", "binning.cu", "blend.cu", "surface.cu", "gaussian_model.cu", "loss.cu", "regularizers.cu", "densify.cu"]
HEADERS = ["common.cuh", "kernels.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "shared"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    deps = [CSRC / s for s in SOURCES + HEADERS] + [ROOT / "include" / "g4s_rasterizer.h"]
    return any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libg4s_rasterizer.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
