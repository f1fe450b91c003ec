"""Golden vectors for compute_mip_filter from the UNMODIFIED reference:
GaussianModel.compute_mip_filter of /root/reference/2d-gaussian-splatting/scene/gaussian_model.py is
imported in this (GPU-less) container and called on CPU tensors.  Modules the container lacks
(simple_knn, plyfile, ...) are stubbed; nothing inside the method is changed.

    python tests/golden/make_golden_mip.py      # writes tests/golden/mip_filter_*.npz
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/2d-gaussian-splatting")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

CASES = {  # name -> (P, cameras, W, H, seed, spread)
    "mip_filter_room": (5000, 12, 320, 200, 21, 3.0),
    "mip_filter_unseen": (800, 3, 96, 64, 22, 30.0),     # most points are outside every frustum
}


def make_case(P, C, W, H, seed, spread):
    """Seeded points and look-at cameras (scene/cameras.py conventions: R is the camera-to-world
    rotation, i.e. xyz_cam = xyz @ R + T)."""
    from g4splat_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    xyz = np.float32(rng.normal(scale=spread, size=(P, 3)))
    cams = []
    for k in range(C):
        ang = 2 * np.pi * k / C
        eye = np.array([2.5 * np.cos(ang), 0.3 * np.sin(3 * ang), 2.5 * np.sin(ang)])
        c = S.look_at_camera(eye, [0.0, 0.0, 0.0], W + 16 * (k % 3), H + 8 * (k % 2), 50.0 + 5 * (k % 4))
        V = np.asarray(c.viewmatrix, dtype=np.float64).reshape(4, 4)     # world_view_transform (row-vector convention)
        R = V[:3, :3].copy()
        T = V[3, :3].copy()
        cams.append(types.SimpleNamespace(R=R, T=T, focal_x=c.W / (2.0 * c.tanfovx), focal_y=c.H / (2.0 * c.tanfovy),
                                          image_width=c.W, image_height=c.H))
    return xyz, cams


def main():
    import make_golden_surface as MS
    sys.path.insert(0, str(REF))
    sys.meta_path.append(MS._StubMissingModules())
    from scene.gaussian_model import GaussianModel      # the reference class, unmodified

    for name, (P, C, W, H, seed, spread) in CASES.items():
        xyz, cams = make_case(P, C, W, H, seed, spread)
        gm = GaussianModel.__new__(GaussianModel)
        gm._xyz = torch.tensor(xyz)
        gm.use_mip_filter = True
        GaussianModel.compute_mip_filter(gm, cams)
        out = gm.mip_filter.numpy()
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", mip_filter=out,
                            meta=np.array(json.dumps({"P": P, "C": C, "W": W, "H": H, "seed": seed, "spread": spread,
                                                      "torch": torch.__version__,
                                                      "reference": "G4Splat scene/gaussian_model.py:388-434, CPU fp32"})))
        print(name, out.shape, "distinct", len(np.unique(out)), "at-max", float((out == out.max()).mean()), flush=True)


if __name__ == "__main__":
    main()
