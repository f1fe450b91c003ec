"""Golden vectors for render()'s post-processing, produced by the UNMODIFIED
reference code: `render()` of /root/reference/2d-gaussian-splatting/gaussian_renderer/__init__.py
(which calls utils/point_utils.py: depth_to_normal) is imported in this (GPU-less) container and run
on the CPU.  Three things are substituted around it, none inside it:
  * modules the container lacks (matplotlib, plyfile, simple_knn, pytorch3d, ...) are stubbed;
  * `GaussianRasterizer` is replaced by a stand-in that returns a prepared (image, radii, allmap)
    -- the block under test starts after that call;
  * `device="cuda"` / `.cuda()` are mapped to the CPU (same torch operators, fp32).

    python tests/golden/make_golden_surface.py          # writes tests/golden/surface_*.npz

Stored per case: the allmap input (CPU oracle forward of a seeded scene with empty pixels), both
camera matrices, depth_ratio, the eight outputs, the upstream gradients' seed and dL_dallmap from
the reference's autograd (NaN where alpha == 0, as the reference produces).
"""
import importlib.abc
import importlib.machinery
import json
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch
import cv2  # noqa: F401  (imported before the stub finder so the real one is used)

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/2d-gaussian-splatting")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

KEYS = ("rend_alpha", "rend_normal", "rend_normal_cam", "rend_dist", "surf_depth", "surf_normal",
        "surf_normal_cam", "rend_depth")
CASES = {  # name -> (P, W, H, seed, splat scale, opacity, depth_ratio)
    "surface_sparse_r1": (60, 80, 48, 11, 0.06, 0.8, 1.0),
    "surface_dense_r03": (400, 72, 56, 12, 0.08, 0.6, 0.3),
}


class _StubMissingModules(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        for f in sys.meta_path:
            if f is self:
                continue
            try:
                if f.find_spec(name, path, target) is not None:
                    return None
            except Exception:
                pass
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__spec__, m.__name__ = [], spec, spec.name
        return m

    def exec_module(self, module):
        pass


def _cpu_for_cuda():
    """torch factory functions: device='cuda' -> CPU; Tensor.cuda() -> identity."""
    patches = [mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self)]
    for name in ("arange", "zeros_like", "ones_like", "tensor", "zeros", "ones"):
        real = getattr(torch, name)

        def wrapped(*a, __real=real, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return __real(*a, **k)
        patches.append(mock.patch.object(torch, name, wrapped))
    return patches


def build_allmap(P, W, H, seed, scale, opacity):
    """allmap of a seeded scene that leaves part of the image empty (CPU oracle forward)."""
    import helpers as Hh
    from g4splat_b200 import synthetic as S
    from oracle.oracle import Oracle
    rng = np.random.default_rng(seed)
    cam = S.look_at_camera([0.3, 0.2, -2.0], [0.0, 0.0, 0.0], W, H, 60.0)
    sc = S.make_scene(P, seed)
    sc["means3D"] = np.float32(rng.uniform([-0.9, -0.6, -0.3], [0.6, 0.6, 1.5], size=(P, 3)))
    sc["scales"] = np.float32(sc["scales"] * 0 + scale * np.exp(rng.normal(scale=0.3, size=(P, 2))))
    sc["opacities"] = np.float32(np.full((P, 1), opacity))
    case = Hh.Case("surface", sc, cam, grad_seed=seed)
    out = Hh.run_oracle(Oracle("f32"), case, backward=False)
    return np.float32(out["allmap"]), cam


def upstream_grads(W, H, seed):
    rng = np.random.default_rng(seed + 1000)
    ch = dict(zip(KEYS, (1, 3, 3, 1, 1, 3, 3, 1)))
    return {k: np.float32(rng.normal(size=(ch[k], H, W)) / (W * H)) for k in KEYS}


def main():
    sys.path.insert(0, str(REF))
    sys.meta_path.append(_StubMissingModules())
    import gaussian_renderer as GR          # the reference module, unmodified

    for name, (P, W, H, seed, scale, opacity, ratio) in CASES.items():
        allmap_np, cam = build_allmap(P, W, H, seed, scale, opacity)
        allmap = torch.tensor(allmap_np, requires_grad=True)
        V, FP = torch.tensor(cam.viewmatrix), torch.tensor(cam.projmatrix)

        class StandInRasterizer:
            def __init__(self, raster_settings):
                pass

            def __call__(self, **kw):
                return torch.zeros(3, H, W), torch.ones(P, dtype=torch.int32), allmap

        view = types.SimpleNamespace(image_width=W, image_height=H, FoVx=cam.FoVx, FoVy=cam.FoVy,
                                     world_view_transform=V, full_proj_transform=FP,
                                     camera_center=torch.tensor(cam.campos), znear=cam.znear, zfar=cam.zfar)
        z = torch.zeros(P, 3)
        pc = types.SimpleNamespace(get_xyz=z, get_opacity=z[:, :1], get_scaling=z[:, :2], get_rotation=z,
                                   get_features=z, active_sh_degree=0, max_sh_degree=3)
        pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, depth_ratio=ratio, debug=False)
        patches = _cpu_for_cuda() + [mock.patch.object(GR, "GaussianRasterizer", StandInRasterizer)]
        for p in patches:
            p.start()
        try:
            pkg = GR.render(view, pc, pipe, torch.zeros(3))
        finally:
            for p in reversed(patches):
                p.stop()
        up = upstream_grads(W, H, seed)
        sum((pkg[k] * torch.tensor(up[k])).sum() for k in KEYS).backward()
        arrays = {k: pkg[k].detach().numpy() for k in KEYS}
        arrays.update(allmap=allmap_np, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                      dL_dallmap=allmap.grad.numpy(),
                      meta=np.array(json.dumps({"W": W, "H": H, "depth_ratio": ratio, "grad_seed": seed,
                                                "torch": torch.__version__,
                                                "reference": "G4Splat ec07361 gaussian_renderer.render + utils.point_utils, CPU fp32"})))
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", **arrays)
        a = arrays["rend_alpha"]
        print(name, "empty pixels:", float((a == 0).mean()), "NaN grads:", int(np.isnan(arrays["dL_dallmap"]).sum()),
              {k: arrays[k].shape for k in KEYS}, flush=True)


if __name__ == "__main__":
    main()
