"""Golden vectors for the photometric loss from the UNMODIFIED reference:
l1_loss and ssim of /root/reference/2d-gaussian-splatting/utils/loss_utils.py are imported in this
(GPU-less) container and combined as train_with_refine_depth.py:382-383 does, on CPU tensors, with
torch autograd for the gradient.

    python tests/golden/make_golden_loss.py      # writes tests/golden/photometric_*.npz
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/2d-gaussian-splatting")

CASES = {  # name -> (C, H, W, seed, lambda_dssim, kind)
    "photometric_small": (3, 40, 56, 41, 0.2, "smooth"),       # smaller than one 32x32 tile row + ragged edges
    "photometric_ragged": (3, 75, 101, 42, 0.2, "noisy"),
    "photometric_ssim_only": (1, 64, 64, 43, 1.0, "smooth"),
}


def make_images(C, H, W, seed, kind):
    """A rendered image and its photograph in [0,1]: smooth structure + texture, the render a slightly
    blurred / shifted / noisier version (flat, saturated and equal regions included)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    gt = np.stack([0.5 + 0.4 * np.sin(6.0 * xx + c) * np.cos(4.0 * yy - c) + 0.1 * rng.normal(size=(H, W)) *
                   (0.2 if kind == "smooth" else 1.0) for c in range(C)])
    gt[:, : H // 4, : W // 4] = 0.0                     # flat black corner (sigma == 0)
    gt[:, -H // 5:, -W // 5:] = 1.0                     # saturated corner
    gt = np.clip(gt, 0.0, 1.0)
    img = np.roll(gt, 1, axis=2) * 0.9 + 0.05 + 0.03 * rng.normal(size=gt.shape)
    img[:, H // 2: H // 2 + 4, :] = gt[:, H // 2: H // 2 + 4, :]      # rows where image == gt exactly (|x-y| kink)
    return np.float32(np.clip(img, 0.0, 1.0)), np.float32(gt)


def main():
    sys.path.insert(0, str(REF))
    from utils.loss_utils import l1_loss, ssim          # the reference functions, unmodified

    for name, (C, H, W, seed, lam, kind) in CASES.items():
        img_np, gt_np = make_images(C, H, W, seed, kind)
        img = torch.tensor(img_np, requires_grad=True)
        gt = torch.tensor(gt_np)
        Ll1 = l1_loss(img, gt)
        s = ssim(img, gt)
        loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)
        loss.backward()
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", loss=loss.detach().numpy(), l1=Ll1.detach().numpy(),
                            ssim=s.detach().numpy(), dL_dimage=img.grad.numpy(),
                            meta=np.array(json.dumps({"C": C, "H": H, "W": W, "seed": seed, "lambda_dssim": lam, "kind": kind,
                                                      "torch": torch.__version__,
                                                      "reference": "G4Splat utils/loss_utils.py l1_loss + ssim, CPU fp32"})))
        print(name, float(loss), float(Ll1), float(s), float(np.abs(img.grad.numpy()).max()), flush=True)


if __name__ == "__main__":
    main()
