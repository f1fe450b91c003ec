"""Build the UNMODIFIED reference rasterizer extension into oracle/_ref/ (git-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  Compiles the five translation units of
  /root/reference/2d-gaussian-splatting/submodules/diff-surfel-rasterization
(cuda_rasterizer/{rasterizer_impl,forward,backward}.cu, rasterize_points.cu, ext.cpp -- the list
in the reference's setup.py:22-28) for sm_100a, from the sources WHERE THEY LIE (nothing is copied
into the repository history), with our own nvcc command lines (the reference's setup.py / CMake
are not run).  The one work-around is `-include cstdint` (gcc 13 vs rasterizer_impl.h:24,40-60).

Outputs (all under oracle/_ref/, which travels to the GPU box with gpurun but is not tracked):
  diff_surfel_rasterization/_C.so      the pybind module the reference's Python wrapper imports
  diff_surfel_rasterization/__init__.py  the reference's own wrapper, installed verbatim (what
                                         `pip install --target` would have placed there)
  BUILD_INFO.json                      compiler, flags, source mtimes

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/2d-gaussian-splatting/submodules/diff-surfel-rasterization")
OUT = HERE / "_ref"
PKG = OUT / "diff_surfel_rasterization"
OBJ = OUT / "obj"
SOURCES = ["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
           "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp"]


def reference_available() -> bool:
    return all((REF / s).exists() for s in SOURCES)


def up_to_date() -> bool:
    so = PKG / "_C.so"
    if not so.exists() or not (PKG / "__init__.py").exists():
        return False
    if not reference_available():
        return True  # GPU box: use the prebuilt files as they are
    newest = max((REF / s).stat().st_mtime for s in SOURCES)
    return so.stat().st_mtime >= newest


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds a usable build."""
    if not force and up_to_date():
        return True
    if not reference_available():
        return False
    import torch  # noqa: F401  (only for the include / library paths)
    from torch.utils.cpp_extension import include_paths, library_paths

    PKG.mkdir(parents=True, exist_ok=True)
    OBJ.mkdir(parents=True, exist_ok=True)
    inc = [f"-I{p}" for p in include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}",
                                                      f"-I{REF / 'third_party/glm'}", f"-I{REF}"]
    common = ["-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1", "-std=c++17", "-O3"]
    nvcc_flags = ["-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                  "-Xcompiler", "-fPIC", "-include", "cstdint", "-w"]
    t0 = time.time()

    def compile_one(src: str) -> str:
        obj = OBJ / (src.replace("/", "_") + ".o")
        if src.endswith(".cu"):
            cmd = ["nvcc", "-c", str(REF / src), "-o", str(obj)] + common + nvcc_flags + inc
        else:
            cmd = ["g++", "-c", str(REF / src), "-o", str(obj), "-fPIC", "-include", "cstdint", "-w"] + common + inc
        if verbose:
            print("[build_ref]", " ".join(cmd[:4]), "...", flush=True)
        subprocess.run(cmd, check=True)
        return str(obj)

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    libdirs = library_paths("cuda")
    link = ["g++", "-shared", "-o", str(PKG / "_C.so")] + objs + [f"-L{p}" for p in libdirs] + \
           ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"] + \
           [f"-Wl,-rpath,{p}" for p in libdirs]
    subprocess.run(link, check=True)
    # "install" step: the reference's own Python wrapper, verbatim, beside its _C module.
    shutil.copyfile(REF / "diff_surfel_rasterization/__init__.py", PKG / "__init__.py")
    info = {"sources": SOURCES, "ref_root": str(REF), "nvcc_flags": nvcc_flags + common,
            "seconds": round(time.time() - t0, 1),
            "nvcc": subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]}
    (OUT / "BUILD_INFO.json").write_text(json.dumps(info, indent=1))
    if verbose:
        print(f"[build_ref] done in {info['seconds']} s -> {PKG / '_C.so'}", flush=True)
    return True


def import_reference():
    """Import the reference package from oracle/_ref under its own name in a private namespace.

    Returns the module object (has GaussianRasterizationSettings, GaussianRasterizer).  Does not
    touch sys.modules['diff_surfel_rasterization'] so it can live beside the product shim."""
    import importlib.util
    import torch  # noqa: F401  must be imported before the extension

    name = "_g4s_reference_dsr"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, PKG / "__init__.py",
                                                  submodule_search_locations=[str(PKG)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no prebuilt files)")
    sys.exit(0 if ok else 1)
