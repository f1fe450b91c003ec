"""The photometric loss: oracle/loss_oracle.py against the golden vectors the
reference's own l1_loss + ssim produced on the CPU, and the fused CUDA kernels against both (`-m gpu`,
through g4s_photometric_forward / g4s_photometric_backward).

Tolerances (floating point): sigma^2 = E[x^2] - mu^2 cancels ~3 digits in flat image regions, so fp32
implementations of ssim differ from each other (and from fp64) by ~1e-4 of the largest gradient --
the reference's own fp32 result sits 1.4e-4 from fp64 on these cases.
  * fp32 oracle vs golden: 1e-6 (same torch operators; conv2d may pick another summation order);
  * kernels vs the fp64 oracle: loss 2e-6 absolute, gradient 3e-4 of its maximum;
  * kernels vs the fp32 golden: loss 2e-6 absolute, gradient 5e-4 of its maximum.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from oracle import loss_oracle as LO  # noqa: E402
import make_golden_loss as MG  # noqa: E402

GOLDEN = sorted((ROOT / "tests" / "golden").glob("photometric_*.npz"))


def _load(path):
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    img, gt = MG.make_images(meta["C"], meta["H"], meta["W"], meta["seed"], meta["kind"])
    return g, meta, img, gt


def _rel(a, b):
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)


def test_golden_files_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_fp32_oracle_reproduces_reference(path):
    g, meta, img, gt = _load(path)
    o = LO.run(img, gt, meta["lambda_dssim"], np.float32)
    for k in ("loss", "l1", "ssim"):
        assert abs(float(o[k]) - float(g[k])) <= 1e-6, k
    assert _rel(o["dL_dimage"], g["dL_dimage"]) <= 1e-6


def test_window_is_the_reference_window():
    import torch
    w = LO.gaussian(11, 1.5)
    assert w.dtype == torch.float32 and abs(float(w.sum()) - 1.0) < 1e-6
    assert np.array_equal(LO.create_window(11, 3)[1, 0].numpy(), np.outer(w.numpy(), w.numpy()).astype(np.float32))


def _run_kernels(img, gt, lam, upstream=1.0):
    import torch
    from g4splat_b200.loss_utils import photometric_loss
    x = torch.tensor(img, device="cuda", requires_grad=True)
    loss, Ll1 = photometric_loss(x, torch.tensor(gt, device="cuda"), lam)
    (loss * upstream).backward()
    return float(loss.detach()), float(Ll1), x.grad.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_kernels_match_reference_and_fp64(path):
    g, meta, img, gt = _load(path)
    lam = meta["lambda_dssim"]
    loss, l1, grad = _run_kernels(img, gt, lam)
    o64 = LO.run(img, gt, lam, np.float64)
    assert abs(loss - float(o64["loss"])) <= 2e-6 and abs(loss - float(g["loss"])) <= 2e-6
    assert abs(l1 - float(o64["l1"])) <= 1e-6
    assert np.isfinite(grad).all()
    assert _rel(grad, o64["dL_dimage"]) <= 3e-4, _rel(grad, o64["dL_dimage"])
    assert _rel(grad, g["dL_dimage"]) <= 5e-4, _rel(grad, g["dL_dimage"])


@pytest.mark.gpu
def test_ssim_wrapper_and_upstream_gradient():
    """ssim() under its reference name; a non-unit upstream gradient scales dL/dimage."""
    import torch
    from g4splat_b200.loss_utils import ssim, l1_loss
    g, meta, img, gt = _load(GOLDEN[0])
    x = torch.tensor(img, device="cuda", requires_grad=True)
    y = torch.tensor(gt, device="cuda")
    s = ssim(x, y)
    o64 = LO.run(img, gt, 1.0, np.float64, upstream=-2.5)     # d(-2.5 * (1 - ssim)) = 2.5 dssim
    (2.5 * s).backward()
    assert abs(float(s.detach()) - float(o64["ssim"])) <= 2e-6
    assert _rel(x.grad.cpu().numpy(), o64["dL_dimage"]) <= 3e-4
    assert abs(float(l1_loss(x, y)) - float(o64["l1"])) <= 1e-6
    with pytest.raises(NotImplementedError):
        ssim(x, y, window_size=7)
    with pytest.raises(RuntimeError):
        ssim(x.detach().cpu(), y.cpu())                       # CPU tensors: no fallback
    with torch.no_grad():                                      # forward only: no derivative maps are allocated
        assert abs(float(ssim(x, y)) - float(o64["ssim"])) <= 2e-6


@pytest.mark.gpu
def test_kernels_full_size_against_fp64():
    """1920x1080x3 (the BASELINE image size): loss and gradient against the fp64 oracle."""
    img, gt = MG.make_images(3, 1080, 1920, 77, "noisy")
    loss, l1, grad = _run_kernels(img, gt, 0.2)
    o64 = LO.run(img, gt, 0.2, np.float64)
    assert abs(loss - float(o64["loss"])) <= 2e-6
    assert _rel(grad, o64["dL_dimage"]) <= 3e-4, _rel(grad, o64["dL_dimage"])
