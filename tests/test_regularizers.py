"""Image-space regularisers: normal2curv and the depth-order loss.

CPU: the torch oracle (oracle/regularizers_oracle.py) against golden vectors produced by the reference's own functions
(tests/golden/make_golden_regularizers.py).  GPU: the fused kernels against the golden vectors and, at 1080p, against
the oracle run on the device.  Tolerances are written next to each assertion."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
CURV = ("regularizers_curv_ones", "regularizers_curv_mask")
ORDER = ("regularizers_order_mean", "regularizers_order_log", "regularizers_order_raw_sum", "regularizers_order_none")


def _meta(g):
    import json
    return json.loads(str(g["meta"]))


@pytest.mark.parametrize("name", CURV)
def test_oracle_normal2curv_matches_reference_golden(name):
    from oracle import regularizers_oracle as RO
    g = np.load(GOLD / f"{name}.npz")
    n = torch.tensor(g["normal"], requires_grad=True)
    curv = RO.normal2curv(n, torch.tensor(g["mask"]))
    (curv * torch.tensor(g["g"])).sum().backward()
    assert np.array_equal(curv.detach().numpy(), g["curv"])          # same torch expressions: bit-equal
    assert np.array_equal(n.grad.numpy(), g["dnormal"])


@pytest.mark.parametrize("name", ORDER)
def test_oracle_depth_order_matches_reference_golden(name):
    from oracle import regularizers_oracle as RO
    g = np.load(GOLD / f"{name}.npz")
    m = _meta(g)
    d = torch.tensor(g["depth"], requires_grad=True)
    loss = RO.depth_order_loss(d, torch.tensor(g["prior"]), torch.tensor(g["shifts"]), m["scene_extent"], m["normalize_loss"],
                               m["log_space"], m["log_scale"], m["reduction"])
    (loss * torch.tensor(g["g"])).sum().backward()
    assert np.allclose(loss.detach().numpy(), g["loss"], rtol=1e-6, atol=1e-9)
    assert np.allclose(d.grad.numpy(), g["ddepth"], rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CURV)
def test_gpu_normal2curv_matches_reference_golden(name):
    from g4splat_b200.regularization import normal2curv
    g = np.load(GOLD / f"{name}.npz")
    n = torch.tensor(g["normal"], device="cuda", requires_grad=True)
    curv = normal2curv(n, torch.tensor(g["mask"], device="cuda"))
    (curv * torch.tensor(g["g"], device="cuda")).sum().backward()
    # forward: the same fp32 operations in the same order, except the final |x| + |y| + |z| association: 2 ulp
    assert np.allclose(curv.detach().cpu().numpy(), g["curv"], rtol=3e-7, atol=1e-7)
    # backward: sums of up to 9 terms in a different order than autograd's: 1e-6 of the largest gradient
    scale = np.abs(g["dnormal"]).max()
    assert np.abs(n.grad.cpu().numpy() - g["dnormal"]).max() <= 1e-6 * scale


@pytest.mark.gpu
def test_gpu_normal2curv_bool_mask_and_1080p():
    from g4splat_b200.regularization import normal2curv
    from oracle import regularizers_oracle as RO
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = torch.randn(3, 1080, 1920, device="cuda", generator=gen)
    mask = torch.rand(1, 1080, 1920, device="cuda", generator=gen) > 0.1
    gup = torch.randn(1, 1080, 1920, device="cuda", generator=gen)
    a = n.clone().requires_grad_(True)
    b = n.clone().requires_grad_(True)
    ca = normal2curv(a, mask)
    cb = RO.normal2curv(b, mask)
    (ca * gup).sum().backward()
    (cb * gup).sum().backward()
    assert float((ca - cb).abs().max()) <= 1e-6 * float(cb.abs().max())
    assert float((a.grad - b.grad).abs().max()) <= 1e-6 * float(b.grad.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ORDER)
def test_gpu_depth_order_matches_reference_golden(name):
    from g4splat_b200 import regularization as R
    g = np.load(GOLD / f"{name}.npz")
    m = _meta(g)
    d = torch.tensor(g["depth"], device="cuda", requires_grad=True)
    shifts = torch.tensor(g["shifts"], device="cuda")
    loss = R._DepthOrder.apply(d, torch.tensor(g["prior"], device="cuda"), shifts, m["scene_extent"], m["normalize_loss"],
                               m["log_space"], m["log_scale"], m["reduction"])
    (loss * torch.tensor(g["g"], device="cuda")).sum().backward()
    # x / extent is x * (1 / extent) here and a division in the CPU golden: 1 ulp per difference, 1e-6 overall
    assert np.allclose(loss.detach().cpu().numpy(), g["loss"], rtol=2e-6, atol=1e-8)
    scale = np.abs(g["ddepth"]).max()
    assert np.abs(d.grad.cpu().numpy() - g["ddepth"]).max() <= 2e-6 * scale


@pytest.mark.gpu
def test_gpu_depth_order_public_function_draws_like_the_reference():
    """compute_depth_order_loss consumes the CUDA generator exactly as the reference's randint call does."""
    from g4splat_b200.regularization import compute_depth_order_loss
    from oracle import regularizers_oracle as RO
    H, W = 1080, 1920
    gen = torch.Generator(device="cuda").manual_seed(9)
    depth = 1.0 + 4.0 * torch.rand(1, H, W, device="cuda", generator=gen)
    prior = depth * 0.8 + 0.3 * torch.randn(1, H, W, device="cuda", generator=gen)
    a = depth.clone().requires_grad_(True)
    b = depth.clone().requires_grad_(True)
    torch.manual_seed(77)
    la = compute_depth_order_loss(a, prior, scene_extent=3.3, max_pixel_shift_ratio=0.05, normalize_loss=True,
                                  log_space=True, log_scale=20., reduction="mean")
    torch.manual_seed(77)
    max_shift = round(0.05 * max(H, W))
    shifts = torch.randint(-max_shift, max_shift + 1, (H * W, 2), device="cuda")     # depth.py:177-178
    lb = RO.depth_order_loss(b, prior, shifts, 3.3, True, True, 20., "mean")
    la.backward()
    lb.backward()
    assert abs(float(la) - float(lb)) <= 2e-6 * abs(float(lb))
    assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())    # colliding atomics: summation order
    with pytest.raises(ValueError):
        compute_depth_order_loss(a, prior, reduction="median")
