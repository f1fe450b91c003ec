"""Image-space regularisers of the reference trainer, fused.

Same names, arguments and return values as the reference functions they replace:

  normal2curv(normal, mask)                 matcha/dm_utils/rendering.py:392-406
  compute_depth_order_loss(depth, prior_depth, scene_extent, max_pixel_shift_ratio, normalize_loss,
                           log_space, log_scale, reduction, debug)
                                            matcha/dm_regularization/depth.py:142-222

so `train_with_refine_depth.py:415,465` runs on them unchanged.  One hand-written kernel per direction each
(csrc/regularizers.cu) instead of 8 / ~15 torch kernels plus autograd.  The depth-order loss draws its random pixel
shifts with the reference's own `torch.randint` call (same shape, same device, same generator), so a seeded run
pairs the same pixels.  No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib

_LIB = _lib.load()


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    return t.to(torch.float32).contiguous()


class _Normal2Curv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normal, mask):
        n = _f32c(normal, "normal")
        if n.dim() != 3 or n.shape[0] != 3:
            raise RuntimeError("normal must have dimensions (3, H, W)")
        H, W = int(n.shape[1]), int(n.shape[2])
        m = None
        if mask is not None:
            m = _f32c(mask, "mask")
            if m.numel() != H * W:
                raise RuntimeError("mask must have dimensions (1, H, W)")
        curv = torch.empty((1, H, W), dtype=torch.float32, device=n.device)
        need_grad = normal.requires_grad
        sg = torch.empty((3, H, W), dtype=torch.float32, device=n.device) if need_grad else None
        with torch.cuda.device(n.device):
            _lib.check(_LIB.g4s_normal2curv_forward(W, H, n.data_ptr(), m.data_ptr() if m is not None else None, curv.data_ptr(),
                                                    sg.data_ptr() if sg is not None else None,
                                                    torch.cuda.current_stream(n.device).cuda_stream))
        ctx.save_for_backward(m if m is not None else n.new_empty(0), sg if sg is not None else n.new_empty(0))
        ctx.has_mask = m is not None
        ctx.size = (W, H)
        return curv

    @staticmethod
    def backward(ctx, g_curv):
        m, sg = ctx.saved_tensors
        W, H = ctx.size
        g = _f32c(g_curv, "dL_dcurv")
        out = torch.empty((3, H, W), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_LIB.g4s_normal2curv_backward(W, H, m.data_ptr() if ctx.has_mask else None, sg.data_ptr(), g.data_ptr(),
                                                     out.data_ptr(), torch.cuda.current_stream(g.device).cuda_stream))
        return out, None


def normal2curv(normal: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """[1,H,W] L1 norm of the masked 4-neighbour Laplacian of a [3,H,W] normal map (replicate padding)."""
    return _Normal2Curv.apply(normal, mask)


class _DepthOrder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, prior_depth, shifts, scene_extent, normalize_loss, log_space, log_scale, reduction):
        d = _f32c(depth, "depth")
        p = _f32c(prior_depth, "prior_depth")
        H, W = (int(s) for s in depth.squeeze().shape)
        N = H * W
        dev = d.device
        per_pixel = torch.empty(depth.shape, dtype=torch.float32, device=dev) if reduction == "none" else None
        total = torch.empty(1, dtype=torch.float64, device=dev) if reduction != "none" else None
        with torch.cuda.device(dev):
            _lib.check(_LIB.g4s_depth_order_forward(W, H, d.data_ptr(), p.data_ptr(), shifts.data_ptr(), float(scene_extent),
                                                    int(bool(normalize_loss)), int(bool(log_space)), float(log_scale),
                                                    per_pixel.data_ptr() if per_pixel is not None else None,
                                                    total.data_ptr() if total is not None else None,
                                                    torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(d, p, shifts)
        ctx.cfg = (W, H, float(scene_extent), int(bool(normalize_loss)), int(bool(log_space)), float(log_scale), reduction, depth.shape)
        if reduction == "none":
            return per_pixel
        out = total / N if reduction == "mean" else total
        return out.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        d, p, shifts = ctx.saved_tensors
        W, H, extent, normalize, log_space, log_scale, reduction, shape = ctx.cfg
        gd = torch.empty((H * W,), dtype=torch.float32, device=d.device)
        g = _f32c(g, "upstream gradient")
        with torch.cuda.device(d.device):
            per = g.data_ptr() if reduction == "none" else None
            scalar = g.data_ptr() if reduction != "none" else None
            scale = 1.0 / (H * W) if reduction == "mean" else 1.0
            _lib.check(_LIB.g4s_depth_order_backward(W, H, d.data_ptr(), p.data_ptr(), shifts.data_ptr(), extent, normalize, log_space,
                                                     log_scale, per, scalar, scale, gd.data_ptr(),
                                                     torch.cuda.current_stream(d.device).cuda_stream))
        return gd.view(shape), None, None, None, None, None, None, None


def compute_depth_order_loss(depth: torch.Tensor, prior_depth: torch.Tensor, scene_extent: float = 1.,
                             max_pixel_shift_ratio: float = 0.05, normalize_loss: bool = True, log_space: bool = False,
                             log_scale: float = 20., reduction: str = "mean", debug: bool = False):
    """Loss encouraging the pixels of `depth` to keep the relative depth order they have in `prior_depth`
    ((H,W), (H,W,1) or (1,H,W) tensors).  Gradient flows to `depth` only (the prior is an input of the trainer)."""
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(f"Invalid reduction: {reduction}")
    if debug:
        raise NotImplementedError("debug=True returns the reference's intermediate tensors; use the reference function for that")
    height, width = depth.squeeze().shape
    # the reference's draw (depth.py:175-178), call for call: same shape, dtype and device -> same random stream
    max_pixel_shift = round(max_pixel_shift_ratio * max(height, width))
    pixel_shifts = torch.randint(-max_pixel_shift, max_pixel_shift + 1, (height * width, 2), device=depth.device)
    return _DepthOrder.apply(depth, prior_depth, pixel_shifts.contiguous(), scene_extent, normalize_loss, log_space, log_scale, reduction)
