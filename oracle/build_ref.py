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
# The reference's own Python call surface around the operator (2d-gaussian-splatting/), installed VERBATIM into
# oracle/_ref/twodgs/ (git-ignored; ships to the GPU box like the .so) so that GPU tests can run the unmodified
# render() / GaussianModel / loss code on either operator.  scene/__init__.py (dataset readers) is not needed.
TWODGS = REF.parent.parent
TWODGS_OUT = OUT / "twodgs"
TWODGS_FILES = ["gaussian_renderer/__init__.py", "scene/gaussian_model.py", "scene/cameras.py", "utils/point_utils.py",
                "utils/sh_utils.py", "utils/general_utils.py", "utils/graphics_utils.py", "utils/loss_utils.py",
                "utils/system_utils.py"]


def reference_available() -> bool:
    return all((REF / s).exists() for s in SOURCES)


def install_twodgs() -> bool:
    """Copy the reference's Python modules (verbatim) next to the extension.  True when they are in place."""
    if all((TWODGS / f).exists() for f in TWODGS_FILES):
        for f in TWODGS_FILES:
            dst = TWODGS_OUT / f
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(TWODGS / f, dst)
    return all((TWODGS_OUT / f).exists() for f in TWODGS_FILES)


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
    install_twodgs()
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


class _StubMissing:
    """Meta-path finder that answers imports nothing else can satisfy (plyfile, simple_knn, cv2, matplotlib: the
    reference imports them at module level, the code paths under test never call them) with MagicMock modules."""

    def find_spec(self, name, path, target=None):
        import importlib.machinery
        if name.split(".")[0] not in ("plyfile", "simple_knn", "cv2", "matplotlib", "open3d", "trimesh"):
            return None
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        from unittest import mock
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__spec__, m.__name__ = [], spec, spec.name
        return m

    def exec_module(self, module):
        pass


def import_twodgs(op_module, tag: str):
    """The reference's gaussian_renderer module (unmodified, from oracle/_ref/twodgs) bound to `op_module` as its
    `diff_surfel_rasterization`.  `tag` names this binding ("b200", "ref"): each gets its own module object, while
    scene.* / utils.* (operator-independent) are shared.  Returns a namespace: render, GaussianModel, Camera, MiniCam,
    l1_loss, ssim, module."""
    import importlib
    import importlib.util
    import types
    if not all((TWODGS_OUT / f).exists() for f in TWODGS_FILES):
        raise ImportError("oracle/_ref/twodgs is missing: run python oracle/build_ref.py where /root/reference exists")
    if not any(isinstance(f, _StubMissing) for f in sys.meta_path):
        sys.meta_path.append(_StubMissing())
    if str(TWODGS_OUT) not in sys.path:
        sys.path.insert(0, str(TWODGS_OUT))
    saved = sys.modules.get("diff_surfel_rasterization")
    sys.modules["diff_surfel_rasterization"] = op_module
    try:
        name = f"_g4s_reference_renderer_{tag}"
        spec = importlib.util.spec_from_file_location(name, TWODGS_OUT / "gaussian_renderer" / "__init__.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["diff_surfel_rasterization"] = saved
        else:
            del sys.modules["diff_surfel_rasterization"]
    gm = importlib.import_module("scene.gaussian_model")
    cams = importlib.import_module("scene.cameras")
    lu = importlib.import_module("utils.loss_utils")
    return types.SimpleNamespace(render=mod.render, GaussianModel=gm.GaussianModel, Camera=cams.Camera, MiniCam=cams.MiniCam,
                                 l1_loss=lu.l1_loss, ssim=lu.ssim, module=mod)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    if reference_available():
        install_twodgs()
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no prebuilt files)")
    sys.exit(0 if ok else 1)
