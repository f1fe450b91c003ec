"""render()'s post-processing of the rasterizer's `allmap`, fused.

Reference: 2d-gaussian-splatting/gaussian_renderer/__init__.py:118-164 (the block after the
rasterizer call) with depth_to_normal / depths_to_points (utils/point_utils.py:9-37).  There it is
~15 torch kernels, an [N,3] ray grid and two matrix inversions per call (torch.inverse synchronises
the host), plus the same graph again in autograd.  Here: one CUDA kernel forward, one backward
(g4s_surface_forward / g4s_surface_backward), no host synchronisation.

Returned keys and shapes are the reference's: rend_alpha[1,H,W], rend_normal[3,H,W],
rend_normal_cam[3,H,W], rend_dist[1,H,W], surf_depth[1,H,W], surf_normal[3,H,W],
surf_normal_cam[3,H,W], rend_depth[1,H,W].
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib

_LIB = _lib.load()  # no fallback: raises when the CUDA library is missing

KEYS = ("rend_alpha", "rend_normal", "rend_normal_cam", "rend_dist", "surf_depth", "surf_normal",
        "surf_normal_cam", "rend_depth")
_CHANNELS = (1, 3, 3, 1, 1, 3, 3, 1)


def _ptr(t):
    return None if t is None else t.data_ptr()


class _SurfaceAttributes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, allmap, viewmatrix, projmatrix, depth_ratio):
        if allmap.dim() != 3 or allmap.shape[0] != 7:
            raise ValueError("allmap must have dimensions (7, H, W)")
        if not allmap.is_cuda:
            raise RuntimeError("allmap must be a CUDA tensor (there is no CPU path)")
        allmap_c = allmap.detach().to(torch.float32).contiguous()
        view_c = viewmatrix.detach().to(device=allmap.device, dtype=torch.float32).contiguous()
        proj_c = projmatrix.detach().to(device=allmap.device, dtype=torch.float32).contiguous()
        H, W = int(allmap.shape[1]), int(allmap.shape[2])
        outs = tuple(torch.empty((c, H, W), dtype=torch.float32, device=allmap.device) for c in _CHANNELS)
        with torch.cuda.device(allmap.device):
            _lib.check(_LIB.g4s_surface_forward(W, H, allmap_c.data_ptr(), view_c.data_ptr(), proj_c.data_ptr(),
                                                float(depth_ratio), *[o.data_ptr() for o in outs],
                                                torch.cuda.current_stream(allmap.device).cuda_stream))
        ctx.save_for_backward(allmap_c, view_c, proj_c)
        ctx.depth_ratio = float(depth_ratio)
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        allmap_c, view_c, proj_c = ctx.saved_tensors
        H, W = int(allmap_c.shape[1]), int(allmap_c.shape[2])
        gs = [None if g is None else g.to(torch.float32).contiguous() for g in grads]
        g_allmap = torch.empty_like(allmap_c)
        with torch.cuda.device(allmap_c.device):
            _lib.check(_LIB.g4s_surface_backward(W, H, allmap_c.data_ptr(), view_c.data_ptr(), proj_c.data_ptr(),
                                                 ctx.depth_ratio, *[_ptr(g) for g in gs], g_allmap.data_ptr(),
                                                 torch.cuda.current_stream(allmap_c.device).cuda_stream))
        return g_allmap, None, None, None


def surface_attributes(allmap: torch.Tensor, world_view_transform: torch.Tensor, full_proj_transform: torch.Tensor,
                       depth_ratio: float) -> Dict[str, torch.Tensor]:
    """The dictionary render() adds to its result (gaussian_renderer/__init__.py:155-164) from the
    rasterizer's allmap and the camera's two matrices.  Differentiable w.r.t. allmap."""
    outs = _SurfaceAttributes.apply(allmap, world_view_transform, full_proj_transform, depth_ratio)
    return dict(zip(KEYS, outs))
