"""Golden vectors for the parameter activations from the UNMODIFIED reference:
the properties get_scaling / get_rotation / get_features / get_opacity of
/root/reference/2d-gaussian-splatting/scene/gaussian_model.py are read on a GaussianModel whose raw
leaves are seeded CPU tensors, with and without the mip filter, and differentiated with torch autograd
against seeded upstream gradients.

    python tests/golden/make_golden_activations.py      # writes tests/golden/activations_*.npz
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/2d-gaussian-splatting")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

RAW_KEYS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")
ACT_KEYS = ("shs", "opacities", "scales", "rotations")
CASES = {"activations_plain": (300, 51, False), "activations_mip": (300, 52, True)}   # name -> (P, seed, mip filter)


def make_raw(P, seed, mip):
    """Raw leaves in the ranges training produces: log-scales around log(0.03), logit opacities,
    un-normalised quaternions (one of them tiny: the eps branch of normalize), SH dc + rest."""
    rng = np.random.default_rng(seed)
    raw = {"_xyz": rng.normal(size=(P, 3)), "_features_dc": rng.normal(scale=1.5, size=(P, 1, 3)),
           "_features_rest": rng.normal(scale=0.1, size=(P, 15, 3)), "_opacity": rng.normal(loc=0.5, scale=2.5, size=(P, 1)),
           "_scaling": np.log(0.03) + rng.normal(scale=0.7, size=(P, 2)), "_rotation": rng.normal(size=(P, 4)) * 1.7}
    raw["_rotation"][0] = 1e-20
    raw = {k: np.float32(v) for k, v in raw.items()}
    mip_filter = np.float32(0.01 + 0.04 * rng.random(size=(P, 1))) if mip else None
    up = {k: np.float32(rng.normal(size=s)) for k, s in
          (("shs", (P, 16, 3)), ("opacities", (P, 1)), ("scales", (P, 2)), ("rotations", (P, 4)))}
    return raw, mip_filter, up


def main():
    import make_golden_surface as MS
    sys.path.insert(0, str(REF))
    sys.meta_path.append(MS._StubMissingModules())
    from scene.gaussian_model import GaussianModel      # the reference class, unmodified

    for name, (P, seed, mip) in CASES.items():
        raw, mip_filter, up = make_raw(P, seed, mip)
        gm = GaussianModel(3)
        for k in RAW_KEYS:
            setattr(gm, k, torch.tensor(raw[k], requires_grad=True))
        gm.use_mip_filter = mip
        if mip:
            gm.mip_filter = torch.tensor(mip_filter)
        act = dict(shs=gm.get_features, opacities=gm.get_opacity, scales=gm.get_scaling, rotations=gm.get_rotation)
        sum((act[k] * torch.tensor(up[k])).sum() for k in ACT_KEYS).backward()
        arrays = {k: act[k].detach().numpy() for k in ACT_KEYS}
        arrays.update({"d" + k: getattr(gm, k).grad.numpy() for k in RAW_KEYS if getattr(gm, k).grad is not None})
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", **arrays,
                            meta=np.array(json.dumps({"P": P, "seed": seed, "mip": mip, "torch": torch.__version__,
                                                      "reference": "G4Splat scene/gaussian_model.py:158-192, CPU fp32"})))
        print(name, {k: v.shape for k, v in arrays.items()}, flush=True)


if __name__ == "__main__":
    main()
