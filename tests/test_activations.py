"""Fused parameter activations: oracle/activations_oracle.py against the golden
vectors read off the reference GaussianModel's own properties (CPU), and the raw-parameter operator
(rasterize_gaussian_model -> g4s_forward_plan_raw / g4s_backward_raw) against "activate with torch,
then call the operator" on the GPU (`-m gpu`).

Tolerances: the activations are a handful of fp32 operations, but exp / sigmoid in the kernel and in
torch's CUDA kernels may differ by an ulp, which can flip an alpha >= 1/255 or T < 1e-4 decision for a
pixel: images use |x - ref| <= 1e-4 * max|ref| + 1e-4 * |ref| with at most 2e-4 of the elements
outside; gradients with respect to the raw leaves use the same form with rtol 1e-3 (they inherit the
atomics' run-to-run noise ~1e-6 and the flipped pixels).
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import helpers as Hh  # noqa: E402
from oracle import activations_oracle as AO  # noqa: E402
import make_golden_activations as MG  # noqa: E402

GOLDEN = sorted((ROOT / "tests" / "golden").glob("activations_*.npz"))


def test_golden_files_exist():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_reproduces_reference_properties(path):
    import torch
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    raw_np, mip_np, up = MG.make_raw(meta["P"], meta["seed"], meta["mip"])
    raw = {k: torch.tensor(v, requires_grad=True) for k, v in raw_np.items()}
    act = AO.activate(raw, None if mip_np is None else torch.tensor(mip_np))
    sum((act[k] * torch.tensor(up[k])).sum() for k in MG.ACT_KEYS).backward()
    for k in MG.ACT_KEYS:
        assert np.array_equal(act[k].detach().numpy(), g[k]), k          # the same torch operators in the same order
    for k in MG.RAW_KEYS:
        if "d" + k in g:
            assert np.array_equal(raw[k].grad.numpy(), g["d" + k]), k


def raw_from_case(case, mip: bool, seed: int):
    """Raw leaves whose activations reproduce the case's scene (up to the mip filter's widening)."""
    rng = np.random.default_rng(seed)
    sc = case.scene
    P = case.P
    op = np.clip(sc["opacities"].astype(np.float64), 1e-4, 1 - 1e-4)
    raw = {"_xyz": sc["means3D"], "_features_dc": sc["shs"][:, :1], "_features_rest": sc["shs"][:, 1:],
           "_opacity": np.log(op / (1 - op)), "_scaling": np.log(sc["scales"].astype(np.float64)),
           "_rotation": sc["rotations"] * rng.uniform(0.3, 3.0, size=(P, 1))}
    raw = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in raw.items()}
    mip_filter = np.float32(sc["scales"].mean() * rng.uniform(0.05, 0.6, size=(P, 1))) if mip else None
    return raw, mip_filter


def _run(case, raw_np, mip_np, fused: bool):
    import torch
    import g4splat_b200.diff_surfel_rasterization as op
    dev = "cuda"
    raw = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in raw_np.items()}
    mip = None if mip_np is None else torch.tensor(mip_np, device=dev)
    means2D = torch.zeros_like(raw["_xyz"], requires_grad=True)
    settings = Hh.make_settings(op, case, dev)
    if fused:
        color, radii, allmap = op.rasterize_gaussian_model(raw["_xyz"], means2D, raw["_features_dc"], raw["_features_rest"],
                                                           raw["_opacity"], raw["_scaling"], raw["_rotation"], mip, settings)
    else:
        act = AO.activate(raw, mip)          # the reference's torch operators, on the GPU
        color, radii, allmap = op.GaussianRasterizer(raster_settings=settings)(
            means3D=act["means3D"], means2D=means2D, opacities=act["opacities"], shs=act["shs"],
            scales=act["scales"], rotations=act["rotations"])
    gc, go = case.upstream()
    ((color * torch.tensor(gc, device=dev)).sum() + (allmap * torch.tensor(go, device=dev)).sum()).backward()
    out = dict(color=color.detach().cpu().numpy(), allmap=allmap.detach().cpu().numpy(), radii=radii.cpu().numpy(),
               dL_dmeans2D=means2D.grad.cpu().numpy())
    out.update({"d" + k: v.grad.cpu().numpy() for k, v in raw.items()})
    torch.cuda.synchronize()
    return out


RAW_GRADS = tuple("d" + k for k in MG.RAW_KEYS) + ("dL_dmeans2D",)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mip", [("tiny", False), ("c0", False), ("c0", True), ("c0_deg1", True), ("ragged", False)])
def test_raw_operator_matches_torch_activations(name, mip, oracle32):
    case = Hh.named_case(name, oracle32)
    if case.colors_precomp is not None or case.transMat_precomp is not None:
        pytest.skip("raw entry points take SH + scale/rotation parameters")
    raw, mip_filter = raw_from_case(case, mip, seed=3)
    want = _run(case, raw, mip_filter, fused=False)
    got = _run(case, raw, mip_filter, fused=True)
    Hh.assert_parity(got, want, ("color", "allmap"), rtol=1e-4, max_bad_frac=2e-4, what=f"{name} mip={mip} forward")
    assert Hh.radii_mismatch(got["radii"], want["radii"], loose=True) <= 2
    Hh.assert_parity(got, want, RAW_GRADS, rtol=1e-3, max_bad_frac=2e-4, what=f"{name} mip={mip} raw gradients")


@pytest.mark.gpu
def test_raw_operator_full_size_and_render_option(oracle32):
    """c1-sized scene (200 k surfels, 1200x680) through render(fused_activations=True) of the reference
    signature, against render() on a model whose properties activate with torch."""
    import types
    import torch
    from g4splat_b200.gaussian_renderer import render
    case = Hh.room_case("act_full", P=200_000, W=1200, H=680, seed=1, cams=5, cam_index=2)
    raw_np, mip_np = raw_from_case(case, True, seed=4)
    cam = case.cam
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")
    view = types.SimpleNamespace(image_width=cam.W, image_height=cam.H, FoVx=cam.FoVx, FoVy=cam.FoVy,
                                 world_view_transform=t(cam.viewmatrix), full_proj_transform=t(cam.projmatrix),
                                 camera_center=t(cam.campos), znear=cam.znear, zfar=cam.zfar)
    pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, depth_ratio=0.0, debug=False)

    class Model:                                   # the attributes render() reads (scene/gaussian_model.py)
        active_sh_degree, max_sh_degree, use_mip_filter = 3, 3, True

        def __init__(self):
            for k, v in raw_np.items():
                setattr(self, k, t(v).requires_grad_(True))
            self.mip_filter = t(mip_np)
        get_xyz = property(lambda s: s._xyz)
        get_scaling = property(lambda s: AO.get_scaling(s._scaling, s.mip_filter))
        get_rotation = property(lambda s: AO.get_rotation(s._rotation))
        get_features = property(lambda s: AO.get_features(s._features_dc, s._features_rest))
        get_opacity = property(lambda s: AO.get_opacity(s._opacity, s._scaling, s.mip_filter))

    res = {}
    for fused in (False, True):
        pc = Model()
        pkg = render(view, pc, pipe, t(case.bg), fused_activations=fused)
        (pkg["render"].mean() + 0.1 * (pkg["rend_normal"] * pkg["surf_normal"]).sum(0).mean() + 0.05 * pkg["rend_dist"].mean()).backward()
        res[fused] = dict(color=pkg["render"].detach().cpu().numpy(), radii=pkg["radii"].cpu().numpy(),
                          **{"d" + k: getattr(pc, k).grad.cpu().numpy() for k in MG.RAW_KEYS})
    Hh.assert_parity(res[True], res[False], ("color",), rtol=1e-4, max_bad_frac=2e-4, what="render forward")
    assert Hh.radii_mismatch(res[True]["radii"], res[False]["radii"], loose=True) <= 4
    Hh.assert_parity(res[True], res[False], tuple("d" + k for k in MG.RAW_KEYS), rtol=1e-3, max_bad_frac=2e-4,
                     what="render raw gradients")


def _truncate_sh(case, M, degree):
    """The same scene with only M allocated SH coefficients (a model trained with max_sh_degree < 3)."""
    import copy
    c = copy.copy(case)
    c.scene = dict(case.scene)
    c.scene["shs"] = np.ascontiguousarray(case.scene["shs"][:, :M])
    c.sh_degree = degree
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("M,degree", [(4, 1), (1, 0), (9, 2)])
def test_fewer_allocated_sh_coefficients(M, degree, b200, oracle32):
    """shs [P,M,3] with M < 16 (reference: M = sh.size(1), rasterize_points.cu:101-105): the operator against
    the CPU oracle, and the raw-parameter operator (features_rest [P,M-1,3], empty for M = 1) against it."""
    case = _truncate_sh(Hh.named_case("tiny", oracle32), M, degree)
    want = Hh.run_oracle(oracle32, case)
    got = Hh.run_operator(b200, case)
    assert got["dL_dsh"].shape == (case.P, M, 3)
    Hh.assert_parity(got, want, ("color", "allmap"), rtol=1e-4, max_bad_frac=2e-4, what=f"M={M} forward vs oracle")
    Hh.assert_parity(got, want, Hh.GRAD_KEYS, rtol=1e-3, max_bad_frac=2e-4, what=f"M={M} backward vs oracle")
    raw, _ = raw_from_case(case, False, seed=5)
    assert raw["_features_rest"].shape == (case.P, M - 1, 3)
    fused = _run(case, raw, None, fused=True)
    Hh.assert_parity(fused, got, ("color", "allmap"), rtol=1e-4, max_bad_frac=2e-4, what=f"M={M} raw forward")
    assert fused["d_features_rest"].shape == (case.P, M - 1, 3)
    np.testing.assert_allclose(fused["d_features_dc"][:, 0], got["dL_dsh"][:, 0], rtol=1e-3, atol=1e-3 * np.abs(got["dL_dsh"]).max())
    if M > 1:
        np.testing.assert_allclose(fused["d_features_rest"], got["dL_dsh"][:, 1:], rtol=1e-3, atol=1e-3 * np.abs(got["dL_dsh"]).max())
