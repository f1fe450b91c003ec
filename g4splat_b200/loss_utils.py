"""The photometric loss of the training loop, fused.

Mirrors 2d-gaussian-splatting/utils/loss_utils.py (l1_loss :17-18, gaussian :29-31, create_window
:44-48, ssim / _ssim :49-80) and the two lines that combine them,
train_with_refine_depth.py:382-383:

    Ll1 = l1_loss(image, gt_image)
    loss = (1.0 - opt.lambda_dssim) * Ll1 + opt.lambda_dssim * (1.0 - ssim(image, gt_image))

    from g4splat_b200.loss_utils import l1_loss, ssim, photometric_loss
    loss, Ll1 = photometric_loss(image, gt_image, opt.lambda_dssim)      # both lines, two kernels

`ssim(img1, img2)` keeps the reference's name and defaults (window_size=11, size_average=True -- the
only configuration the reference uses); both run g4s_photometric_forward / _backward.  No CPU path.
"""
from __future__ import annotations

from math import exp

import torch

from . import _lib

_LIB = _lib.load()

WINDOW_SIZE = 11


def gaussian(window_size: int = WINDOW_SIZE, sigma: float = 1.5) -> torch.Tensor:
    """loss_utils.py:29-31, the same expression (fp32 sum and division), on the CPU."""
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


_WINDOW = gaussian().contiguous()
_WINDOW_C = (_lib.C.c_float * WINDOW_SIZE)(*[float(v) for v in _WINDOW])


def _as_chw(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3:
        raise ValueError(f"{what} must have dimensions (C, H, W) or (1, C, H, W)")
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor (there is no CPU path)")
    return t


class _Photometric(torch.autograd.Function):
    """(image, gt, lambda) -> out3 = [loss, l1, ssim]; only out3[0] carries a gradient."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        img = image.detach().to(torch.float32).contiguous()
        ref = gt.detach().to(device=img.device, dtype=torch.float32).contiguous()
        if img.shape != ref.shape:
            raise ValueError("image and gt must have the same shape")
        C, H, W = (int(v) for v in img.shape)
        need_grad = ctx.needs_input_grad[0]
        out3 = torch.empty(3, dtype=torch.float32, device=img.device)
        sums = torch.empty(2, dtype=torch.float64, device=img.device)
        dmaps = torch.empty((3, C, H, W), dtype=torch.float32, device=img.device) if need_grad else None
        with torch.cuda.device(img.device):
            _lib.check(_LIB.g4s_photometric_forward(W, H, C, img.data_ptr(), ref.data_ptr(), _WINDOW_C, float(lambda_dssim),
                                                    sums.data_ptr(), None if dmaps is None else dmaps.data_ptr(),
                                                    out3.data_ptr(), torch.cuda.current_stream(img.device).cuda_stream))
        if need_grad:
            ctx.save_for_backward(img, ref, dmaps)
        ctx.lambda_dssim = float(lambda_dssim)
        return out3

    @staticmethod
    def backward(ctx, g_out3):
        img, ref, dmaps = ctx.saved_tensors
        C, H, W = (int(v) for v in img.shape)
        g = g_out3[0:1].to(torch.float32).contiguous()      # d/d loss; l1 and ssim entries are reported values only
        grad = torch.empty_like(img)
        with torch.cuda.device(img.device):
            _lib.check(_LIB.g4s_photometric_backward(W, H, C, img.data_ptr(), ref.data_ptr(), _WINDOW_C, ctx.lambda_dssim,
                                                     dmaps.data_ptr(), g.data_ptr(), grad.data_ptr(),
                                                     torch.cuda.current_stream(img.device).cuda_stream))
        return grad, None, None


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2):
    """(loss, Ll1) of train_with_refine_depth.py:382-383; Ll1 is detached (the trainer only logs it)."""
    out3 = _Photometric.apply(_as_chw(image, "image"), _as_chw(gt, "gt"), lambda_dssim)
    return out3[0], out3[1].detach()


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:17-18 (one torch expression there too; kept for drop-in imports)."""
    return torch.abs(network_output - gt).mean()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = WINDOW_SIZE, size_average: bool = True) -> torch.Tensor:
    """loss_utils.py:49-80 for the configuration the reference uses."""
    if window_size != WINDOW_SIZE or not size_average:
        raise NotImplementedError("fused ssim supports window_size=11, size_average=True (the reference's only use)")
    # loss at lambda = 1 is 1 - ssim
    return 1.0 - _Photometric.apply(_as_chw(img1, "img1"), _as_chw(img2, "img2"), 1.0)[0]
