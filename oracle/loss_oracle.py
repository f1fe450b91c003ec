"""TEST INFRASTRUCTURE -- CPU restatement (torch, fp32 or fp64) of the trainer's photometric loss.

Only tests/ and tests/tools/ may import this; the product path (g4splat_b200/loss_utils.py ->
g4s_photometric_forward/backward) never does.  Follows
  2d-gaussian-splatting/utils/loss_utils.py:17-18    l1_loss
  2d-gaussian-splatting/utils/loss_utils.py:29-31    gaussian
  2d-gaussian-splatting/utils/loss_utils.py:44-48    create_window
  2d-gaussian-splatting/utils/loss_utils.py:49-80    ssim, _ssim
  2d-gaussian-splatting/train_with_refine_depth.py:382-383   the combination
Pinned by tests/golden/photometric_*.npz, which tests/golden/make_golden_loss.py produced by calling
the reference's own functions on the CPU.  Gradients come from torch autograd, as in the reference.
"""
from __future__ import annotations

from math import exp

import numpy as np
import torch
import torch.nn.functional as F


def gaussian(window_size, sigma):
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    channel = img1.size(-3)
    window = create_window(window_size, channel).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def run(image: np.ndarray, gt: np.ndarray, lambda_dssim: float, dtype=np.float32, upstream: float = 1.0):
    """loss, Ll1, ssim and dloss/dimage (times `upstream`) as numpy."""
    td = torch.float32 if dtype == np.float32 else torch.float64
    img = torch.tensor(np.asarray(image), dtype=td, requires_grad=True)
    ref = torch.tensor(np.asarray(gt), dtype=td)
    Ll1 = l1_loss(img, ref)
    s = ssim(img, ref)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s)
    (loss * upstream).backward()
    return dict(loss=loss.detach().numpy(), l1=Ll1.detach().numpy(), ssim=s.detach().numpy(), dL_dimage=img.grad.numpy())
