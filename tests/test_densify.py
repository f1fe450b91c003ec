"""Adaptive density control: GaussianModel.densify_and_prune, fused.

CPU: the torch restatement (oracle/densify_oracle.py) against golden vectors produced by the reference's own
GaussianModel.densify_and_prune + Adam optimizer (tests/golden/make_golden_densify.py).  GPU: the two fused kernels behind
g4splat_b200.gaussian_model.densify_and_prune against the same golden vectors (same split samples handed in), the
replica-identical seeded variant, and the re-binding of the view-sharded gradient buffer when P changes."""
import json
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
CASES = ("densify_plain", "densify_screen")
KEYS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
        "rotation": "_rotation"}


def _load(name):
    g = np.load(GOLD / f"{name}.npz")
    return g, json.loads(str(g["meta"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import densify_oracle as DO
    g, m = _load(name)
    params = {k: torch.tensor(g[f"in_{k}"]) for k in KEYS}
    moments = {k: (torch.tensor(g[f"in_m1_{k}"]), torch.tensor(g[f"in_m2_{k}"])) for k in KEYS}
    p, mo = DO.densify_and_prune(params, moments, torch.tensor(g["accum"]), torch.tensor(g["denom"]), m["percent_dense"],
                                 m["max_grad"], m["min_opacity"], m["extent"], m["max_screen_size"], torch.tensor(g["samples"]))
    assert p["xyz"].shape[0] == m["P_out"]
    for k in KEYS:      # the same torch expressions in the same order: bit-equal
        assert np.array_equal(p[k].numpy(), g[f"out_{k}"]), k
        assert np.array_equal(mo[k][0].numpy(), g[f"out_m1_{k}"]) and np.array_equal(mo[k][1].numpy(), g[f"out_m2_{k}"]), k


def _gpu_model(g, m, device="cuda"):
    """Duck-typed GaussianModel + the Adam optimizer of training_setup, with the golden's moments and step counters."""
    model = types.SimpleNamespace(percent_dense=m["percent_dense"])
    groups = []
    for k in KEYS:
        p = torch.nn.Parameter(torch.tensor(g[f"in_{k}"], device=device).requires_grad_(True))
        setattr(model, ATTR[k], p)
        groups.append({"params": [p], "lr": 1e-3, "name": k})
    model.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for k in KEYS:
        p = getattr(model, ATTR[k])
        model.optimizer.state[p] = {"step": torch.tensor(m["steps"][k]), "exp_avg": torch.tensor(g[f"in_m1_{k}"], device=device),
                                    "exp_avg_sq": torch.tensor(g[f"in_m2_{k}"], device=device)}
    model.xyz_gradient_accum = torch.tensor(g["accum"], device=device)
    model.denom = torch.tensor(g["denom"], device=device)
    model.max_radii2D = torch.tensor(g["max_radii"], device=device)
    return model


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_densify_and_prune_matches_reference_golden(name):
    from g4splat_b200.gaussian_model import densify_and_prune
    g, m = _load(name)
    model = _gpu_model(g, m)
    P_new = densify_and_prune(model, m["max_grad"], m["min_opacity"], m["extent"], m["max_screen_size"],
                              _samples=torch.tensor(g["samples"]))
    assert P_new == m["P_out"]
    for k in KEYS:
        p = getattr(model, ATTR[k])
        assert isinstance(p, torch.nn.Parameter) and p.requires_grad and model.optimizer.param_groups[KEYS.index(k)]["params"][0] is p
        got, want = p.detach().cpu().numpy(), g[f"out_{k}"]
        assert got.shape == want.shape, k
        if k in ("xyz", "scaling"):
            # children: R(q) sample + xyz and log(exp(s) / 1.6) in fp32 on the device vs torch on the CPU: 1e-6 of the values' scale
            assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), k
        else:
            assert np.array_equal(got, want), k                       # pure row copies
        st = model.optimizer.state[p]
        assert np.array_equal(st["exp_avg"].cpu().numpy(), g[f"out_m1_{k}"]), k
        assert np.array_equal(st["exp_avg_sq"].cpu().numpy(), g[f"out_m2_{k}"]), k
        assert float(st["step"]) == m["steps"][k]
    assert tuple(model.xyz_gradient_accum.shape) == (P_new, 1) and float(model.xyz_gradient_accum.abs().sum()) == 0.0
    assert tuple(model.denom.shape) == (P_new, 1) and tuple(model.max_radii2D.shape) == (P_new,)
    model.optimizer.step()      # the rebuilt optimizer is usable (no gradients yet: a no-op, but it walks every group)


@pytest.mark.gpu
def test_gpu_seeded_densification_is_replica_identical_and_rebinds_the_gradient_buffer():
    """Two 'ranks' (two models on one device) with identically seeded generators densify identically, and the
    view-sharded gradient buffer follows the new parameter tensors (P changes, P % 4 != 0 afterwards)."""
    import g4splat_b200.diff_surfel_rasterization as op
    from g4splat_b200.gaussian_model import densify_and_prune
    from g4splat_b200.view_parallel import ViewShardedGradSync
    g, m = _load("densify_plain")
    models = [_gpu_model(g, m), _gpu_model(g, m)]
    syncs = []
    for model in models:
        params = {k: getattr(model, ATTR[k]) for k in KEYS}
        s = ViewShardedGradSync(params, transport="nccl")
        syncs.append(s)
    syncs[0].bind(op, names=["xyz", "opacity", "scaling", "rotation"])
    for model, s in zip(models, syncs):
        gen = torch.Generator(device="cuda").manual_seed(1234)
        P_new = densify_and_prune(model, m["max_grad"], m["min_opacity"], m["extent"], m["max_screen_size"], generator=gen)
        s.rebind({k: getattr(model, ATTR[k]) for k in KEYS})
        assert s.P == P_new and s.flat.numel() >= P_new * 58
        for k in KEYS:
            p = getattr(model, ATTR[k])
            assert p.grad is not None and p.grad.shape == p.shape and p.grad.data_ptr() == s._views[k].data_ptr()
    for k in KEYS:
        assert torch.equal(getattr(models[0], ATTR[k]), getattr(models[1], ATTR[k])), k
    op.set_gradient_sink(None)
