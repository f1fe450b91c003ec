"""Golden vectors for adaptive density control from the UNMODIFIED reference:
GaussianModel.densify_and_prune of /root/reference/2d-gaussian-splatting/scene/gaussian_model.py (with the Adam optimizer
that training_setup builds, after a few steps so that the moments are non-trivial) is run on the CPU -- the file's
hard-coded device="cuda" arguments are redirected to the CPU for the duration -- and the torch.normal draw of
densify_and_split is recorded so that the code under test can be handed the same samples.

    python tests/golden/make_golden_densify.py      # writes tests/golden/densify_*.npz
"""
import json
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/2d-gaussian-splatting")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

CASES = {"densify_plain": dict(P=700, seed=61, max_screen_size=None, min_opacity=0.05),
         "densify_screen": dict(P=500, seed=62, max_screen_size=20, min_opacity=0.005)}
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
        "rotation": "_rotation"}


def make_inputs(P, seed):
    rng = np.random.default_rng(seed)
    raw = {"xyz": rng.normal(size=(P, 3)), "f_dc": rng.normal(size=(P, 1, 3)), "f_rest": rng.normal(scale=0.1, size=(P, 15, 3)),
           "opacity": rng.normal(loc=-1.0, scale=2.5, size=(P, 1)), "scaling": np.log(0.05) + rng.normal(scale=1.2, size=(P, 2)),
           "rotation": rng.normal(size=(P, 4))}
    raw = {k: np.float32(v) for k, v in raw.items()}
    accum = np.float32(np.abs(rng.normal(scale=0.002, size=(P, 1))) * rng.integers(0, 12, size=(P, 1)))
    denom = np.float32(rng.integers(0, 12, size=(P, 1)))          # zeros in the denominator: NaN gradients -> 0
    max_radii = np.float32(rng.integers(0, 60, size=(P,)))
    return raw, accum, denom, max_radii


def main():
    import make_golden_surface as MS
    from oracle import build_ref
    # verbatim copies of the reference's files (scene/__init__.py, which pulls in the dataset readers, is left out)
    assert build_ref.reference_available() and build_ref.install_twodgs()
    sys.path.insert(0, str(build_ref.TWODGS_OUT))
    sys.meta_path.append(build_ref._StubMissing())      # answers only plyfile / simple_knn / cv2 / matplotlib
    from scene.gaussian_model import GaussianModel      # the reference class, unmodified
    opt_args = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                     position_lr_delay_mult=0.01, position_lr_max_steps=30_000, feature_lr=0.0025,
                                     opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    for name, c in CASES.items():
        raw, accum, denom, max_radii = make_inputs(c["P"], c["seed"])
        patches = MS._cpu_for_cuda() + [mock.patch.object(torch.cuda, "empty_cache", lambda: None)]
        for p in patches:
            p.start()
        try:
            gm = GaussianModel(3)
            for k, attr in ATTR.items():
                setattr(gm, attr, torch.nn.Parameter(torch.tensor(raw[k]).requires_grad_(True)))
            gm.spatial_lr_scale = 4.0
            gm.training_setup(opt_args)
            # three Adam steps on seeded gradients: non-trivial moments and step counters
            g = torch.Generator().manual_seed(c["seed"])
            for _ in range(3):
                for attr in ATTR.values():
                    t = getattr(gm, attr)
                    t.grad = torch.randn(t.shape, generator=g) * 1e-3
                gm.optimizer.step()
            before = {k: getattr(gm, a).detach().clone().numpy() for k, a in ATTR.items()}
            moments = {k: (gm.optimizer.state[getattr(gm, a)]["exp_avg"].clone().numpy(),
                           gm.optimizer.state[getattr(gm, a)]["exp_avg_sq"].clone().numpy()) for k, a in ATTR.items()}
            gm.xyz_gradient_accum = torch.tensor(accum)
            gm.denom = torch.tensor(denom)
            gm.max_radii2D = torch.tensor(max_radii)
            drawn = []
            real_normal = torch.normal

            def recording_normal(*a, **k):
                t = real_normal(*a, **k)
                drawn.append(t.clone())
                return t

            torch.manual_seed(c["seed"])
            extent = 5.0
            with mock.patch.object(torch, "normal", recording_normal):
                gm.densify_and_prune(0.0002, c["min_opacity"], extent, c["max_screen_size"])
            assert len(drawn) == 1
            after = {k: getattr(gm, a).detach().numpy() for k, a in ATTR.items()}
            after_m = {k: (gm.optimizer.state[getattr(gm, a)]["exp_avg"].numpy(), gm.optimizer.state[getattr(gm, a)]["exp_avg_sq"].numpy())
                       for k, a in ATTR.items()}
            steps = {k: float(gm.optimizer.state[getattr(gm, a)]["step"]) for k, a in ATTR.items()}
        finally:
            for p in reversed(patches):
                p.stop()
        arrays = {"accum": accum, "denom": denom, "max_radii": max_radii, "samples": drawn[0].detach().numpy()}
        for k in ATTR:
            arrays[f"in_{k}"] = before[k]; arrays[f"in_m1_{k}"] = moments[k][0]; arrays[f"in_m2_{k}"] = moments[k][1]
            arrays[f"out_{k}"] = after[k]; arrays[f"out_m1_{k}"] = after_m[k][0]; arrays[f"out_m2_{k}"] = after_m[k][1]
        meta = dict(c, extent=extent, max_grad=0.0002, percent_dense=0.01, steps=steps, P_out=int(after["xyz"].shape[0]),
                    stats_after=[int(gm.xyz_gradient_accum.shape[0]), float(gm.xyz_gradient_accum.abs().sum()), float(gm.max_radii2D.abs().sum())],
                    reference="G4Splat scene/gaussian_model.py:528-647, CPU fp32", torch=torch.__version__)
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", **arrays, meta=np.array(json.dumps(meta)))
        print(name, "P", c["P"], "->", meta["P_out"], "split samples", drawn[0].shape)


if __name__ == "__main__":
    main()
