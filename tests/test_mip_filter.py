"""compute_mip_filter: the numpy oracle against the golden vectors the
reference's own GaussianModel.compute_mip_filter produced on the CPU, and the fused CUDA kernels
(g4s_mip_filter, through g4splat_b200.gaussian_model.compute_mip_filter) against both (`-m gpu`).

Tolerance: the filter is distance / focal * sqrt(variance) with distance one clamped camera-space
depth; the only arithmetic that can differ from the reference is the summation order inside
`xyz @ R` (BLAS), i.e. a few ulp of z: |x - ref| <= 2e-6 * |ref| element-wise.  A point whose
projection lies within those ulps of the enlarged image border may be seen by one implementation and
not the other; such points are counted and must stay below 1e-3 of P.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from oracle import mip_filter_oracle as MO  # noqa: E402
import make_golden_mip as MG  # noqa: E402

GOLDEN = sorted((ROOT / "tests" / "golden").glob("mip_filter_*.npz"))
RTOL, FLIP_BUDGET = 2e-6, 1e-3


def _load(path):
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    xyz, cams = MG.make_case(meta["P"], meta["C"], meta["W"], meta["H"], meta["seed"], meta["spread"])
    return g["mip_filter"], xyz, cams


def _check(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    bad = np.abs(got - want) > RTOL * np.abs(want)
    assert bad.mean() <= FLIP_BUDGET, (int(bad.sum()), got[bad][:5], want[bad][:5])


def test_golden_files_exist():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_reproduces_reference(path):
    want, xyz, cams = _load(path)
    _check(MO.compute_mip_filter(xyz, cams), want)


def test_oracle_unseen_points_take_the_largest_distance():
    want, xyz, cams = _load(GOLDEN[-1])
    got = MO.compute_mip_filter(xyz, cams)
    assert (got == got.max()).mean() > 0.5      # most points of this case are outside every frustum


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_kernel_matches_reference_and_oracle(path):
    import torch
    from g4splat_b200.gaussian_model import compute_mip_filter
    want, xyz, cams = _load(path)
    got = compute_mip_filter(torch.tensor(xyz, device="cuda"), cams).cpu().numpy()
    _check(got, want)
    _check(got, MO.compute_mip_filter(xyz, cams))


@pytest.mark.gpu
def test_kernel_many_cameras_and_arguments():
    """More cameras than one shared-memory chunk (256), non-default znear / variance."""
    import torch
    from g4splat_b200.gaussian_model import compute_mip_filter
    xyz, cams = MG.make_case(20_000, 300, 200, 120, 31, 2.0)
    got = compute_mip_filter(torch.tensor(xyz, device="cuda"), cams, znear=0.5, filter_variance=0.35).cpu().numpy()
    _check(got, MO.compute_mip_filter(xyz, cams, znear=0.5, filter_variance=0.35))


@pytest.mark.gpu
def test_kernel_error_behaviour():
    import torch
    from g4splat_b200.gaussian_model import compute_mip_filter
    xyz, cams = MG.make_case(64, 2, 64, 48, 5, 1.0)
    far = torch.tensor(xyz, device="cuda") + 1.0e4         # nothing in any frustum: the reference raises too
    with pytest.raises(RuntimeError):
        compute_mip_filter(far, cams)
    with pytest.raises(RuntimeError):
        compute_mip_filter(torch.tensor(xyz), cams)        # CPU tensor: no fallback
    assert compute_mip_filter(torch.zeros((0, 3), device="cuda"), cams).shape == (0, 1)
