"""TEST INFRASTRUCTURE -- CPU restatement of render()'s post-processing of `allmap`.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this; the product
path (g4splat_b200/surface.py -> g4s_surface_forward/backward) never does.

Follows, in torch on the CPU (fp32 = the reference's arithmetic, fp64 = the yardstick):
  2d-gaussian-splatting/gaussian_renderer/__init__.py:118-164   the block after the rasterizer call
  2d-gaussian-splatting/utils/point_utils.py:9-24               depths_to_points
  2d-gaussian-splatting/utils/point_utils.py:26-37              depth_to_normal
Pinned by tests/golden/surface_*.npz, which tests/golden/make_golden_surface.py produced by running
the reference's own render() (CPU, stub rasterizer) on the same inputs.
Gradients come from torch autograd over the same graph, as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

KEYS = ("rend_alpha", "rend_normal", "rend_normal_cam", "rend_dist", "surf_depth", "surf_normal",
        "surf_normal_cam", "rend_depth")


def pixel_rays(V: torch.Tensor, FP: torch.Tensor, W: int, H: int):
    """point_utils.py:10-22: world-space ray directions [H*W,3] and the ray origin [3]."""
    dt, dev = V.dtype, V.device
    c2w = V.T.inverse()
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=dt, device=dev).T
    intrins = ((c2w.T @ FP) @ ndc2pix)[:3, :3].T
    gx, gy = torch.meshgrid(torch.arange(W, dtype=dt, device=dev), torch.arange(H, dtype=dt, device=dev), indexing="xy")
    pix = torch.stack([gx, gy, torch.ones_like(gx)], dim=-1).reshape(-1, 3)
    return pix @ intrins.inverse().T @ c2w[:3, :3].T, c2w[:3, 3]


def normals_from_depth(V, FP, depth):
    """point_utils.py:26-37: central differences of the back-projected depth map, [H,W,3]."""
    H, W = depth.shape[-2:]
    rays_d, rays_o = pixel_rays(V, FP, W, H)
    pts = (depth.reshape(-1, 1) * rays_d + rays_o).reshape(H, W, 3)
    out = torch.zeros_like(pts)
    d_rows = pts[2:, 1:-1] - pts[:-2, 1:-1]
    d_cols = pts[1:-1, 2:] - pts[1:-1, :-2]
    out[1:-1, 1:-1, :] = torch.nn.functional.normalize(torch.cross(d_rows, d_cols, dim=-1), dim=-1)
    return out


def surface_attributes(allmap, V, FP, depth_ratio: float):
    """gaussian_renderer/__init__.py:118-164 on torch tensors (any float dtype; CPU in the tests, or the
    GPU when tests/tools/bench_surface.py times the reference's operator sequence)."""
    alpha = allmap[1:2]
    n_cam = allmap[2:5]
    rend_normal = (n_cam.permute(1, 2, 0) @ V[:3, :3].T).permute(2, 0, 1)
    median = torch.nan_to_num(allmap[5:6], 0, 0)
    expected = torch.nan_to_num(allmap[0:1] / alpha, 0, 0)
    surf_depth = expected * (1 - depth_ratio) + depth_ratio * median
    surf_normal = normals_from_depth(V, FP, surf_depth).permute(2, 0, 1) * alpha.detach()
    surf_normal_cam = (surf_normal.permute(1, 2, 0) @ V[:3, :3]).permute(2, 0, 1)
    return dict(rend_alpha=alpha, rend_normal=rend_normal, rend_normal_cam=n_cam.clone(), rend_dist=allmap[6:7],
                surf_depth=surf_depth, surf_normal=surf_normal, surf_normal_cam=surf_normal_cam, rend_depth=expected)


def run(allmap: np.ndarray, view: np.ndarray, proj: np.ndarray, depth_ratio: float, upstream=None, dtype=np.float32):
    """numpy in / numpy out.  upstream: dict key -> gradient array (missing keys = no gradient);
    when given, the result also holds dL_dallmap [7,H,W] (NaN where the reference's autograd
    produces NaN: alpha == 0)."""
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    am = torch.tensor(np.asarray(allmap), dtype=tdt, requires_grad=upstream is not None)
    V, FP = torch.tensor(np.asarray(view), dtype=tdt), torch.tensor(np.asarray(proj), dtype=tdt)
    out = surface_attributes(am, V, FP, depth_ratio)
    res = {k: v.detach().numpy().copy() for k, v in out.items()}
    if upstream is not None:
        loss = sum((out[k] * torch.tensor(np.asarray(g), dtype=tdt)).sum() for k, g in upstream.items())
        loss.backward()
        res["dL_dallmap"] = am.grad.numpy().copy()
    return res
