"""Per-Gaussian bookkeeping the trainer runs around the rasterizer, fused.

`compute_mip_filter(xyz, cameras, znear, filter_variance)` mirrors
GaussianModel.compute_mip_filter (2d-gaussian-splatting/scene/gaussian_model.py:388-434): same
arguments and meaning, returns the [P,1] tensor the reference stores in `self.mip_filter`.  Cameras are
the reference's Camera objects, duck-typed: R, T (numpy 3x3 / 3), focal_x, focal_y, image_width,
image_height (scene/cameras.py:27-28,39-40,63-66).

The reference runs ~14 torch kernels per camera on [P]-sized tensors; here it is two launches for
any number of cameras (g4s_mip_filter).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_LIB = _lib.load()

CAMERA_RECORD_FLOATS = 20


def camera_records(cameras) -> np.ndarray:
    """[C,20] fp32 records of include/g4s_rasterizer.h: every Python-scalar product of the
    reference (W / 2.0, -0.15 * W, ...) is formed in double and rounded once, as torch does when it
    combines a Python scalar with a float32 tensor."""
    rec = np.zeros((len(cameras), CAMERA_RECORD_FLOATS), dtype=np.float64)
    for i, cam in enumerate(cameras):
        W, H = float(cam.image_width), float(cam.image_height)
        rec[i, 0:9] = np.asarray(cam.R, dtype=np.float32).reshape(9)
        rec[i, 9:12] = np.asarray(cam.T, dtype=np.float32).reshape(3)
        rec[i, 12:] = (cam.focal_x, cam.focal_y, W / 2.0, H / 2.0, -0.15 * W, W * 1.15, -0.15 * H, 1.15 * H)
    return rec.astype(np.float32)


def compute_mip_filter(xyz: torch.Tensor, cameras, znear: float = 0.2, filter_variance: float = 0.2) -> torch.Tensor:
    if not xyz.is_cuda:
        raise RuntimeError("xyz must be a CUDA tensor (there is no CPU path)")
    if xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must have dimensions (num_points, 3)")
    cameras = list(cameras)
    P = int(xyz.shape[0])
    xyz_c = xyz.detach().to(torch.float32).contiguous()
    focal_length = 0.0
    for cam in cameras:                                   # :428-429
        if focal_length < cam.focal_x:
            focal_length = cam.focal_x
    if not cameras or not focal_length > 0.0:
        raise RuntimeError("compute_mip_filter needs at least one camera with a positive focal length")
    rec = torch.from_numpy(camera_records(cameras)).to(xyz.device)
    out = torch.empty((P, 1), dtype=torch.float32, device=xyz.device)
    max_bits = torch.empty(1, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_LIB.g4s_mip_filter(P, xyz_c.data_ptr(), len(cameras), rec.data_ptr(), float(znear),
                                       float(focal_length), float(filter_variance ** 0.5), out.data_ptr(),
                                       max_bits.data_ptr(), torch.cuda.current_stream(xyz.device).cuda_stream))
    if P > 0 and int(max_bits.item()) == 0:
        # the reference fails here too: distance[valid_points].max() of an empty selection (:431)
        raise RuntimeError("compute_mip_filter: no point is seen by any camera")
    return out


# ---- adaptive density control ---------------------------------------------------------------------
_GROUPS = (("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"),
           ("scaling", "_scaling"), ("rotation", "_rotation"))


def densify_and_prune(model, max_grad, min_opacity, extent, max_screen_size, generator=None, N: int = 2, _samples=None):
    """GaussianModel.densify_and_prune (2d-gaussian-splatting/scene/gaussian_model.py:621-640), fused.

    `model` is the reference's GaussianModel (duck-typed: _xyz, _features_dc, _features_rest, _opacity, _scaling,
    _rotation, xyz_gradient_accum, denom, max_radii2D, percent_dense, optimizer with one parameter per group named
    xyz / f_dc / f_rest / opacity / scaling / rotation).  Same arguments, same effect on the model and on the optimizer
    (new nn.Parameters, Adam moments carried for survivors and zero for new points, `step` kept, statistics reset), same
    final row order [survivors | clones | first split children | second split children].  Clone / split / prune are
    decided by one kernel and applied by one gather kernel instead of ~60 torch kernels.

    The split samples are drawn with the reference's own call `torch.normal(mean=zeros[2S,3], std=stds)` on the model's
    device, so a seeded run creates the same children.  `generator`: a torch.Generator for that draw -- give every rank
    of a view-sharded run an identically seeded one and the replicas stay identical; call
    ViewShardedGradSync.rebind(new parameters) afterwards.  (As in the reference, the mip filter plays no part here:
    :622-624 switch it off for the duration.)  Returns the number of Gaussians after the update."""
    import torch.nn as nn
    xyz = model._xyz
    if not xyz.is_cuda:
        raise RuntimeError("densify_and_prune needs CUDA tensors (there is no CPU path)")
    dev = xyz.device
    P = int(xyz.shape[0])
    f32c = lambda t: t.detach().to(torch.float32).contiguous()
    old = [f32c(getattr(model, attr)) for _, attr in _GROUPS]
    rest_w = int(old[2].numel() // max(P, 1))
    flags = torch.empty((P,), dtype=torch.uint8, device=dev)
    accum, denom = f32c(model.xyz_gradient_accum).view(-1), f32c(model.denom).view(-1)
    big_ws = 0.1 * float(extent) if max_screen_size else -1.0
    with torch.cuda.device(dev):
        sp = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_LIB.g4s_densify_classify(P, accum.data_ptr(), denom.data_ptr(), old[4].data_ptr(), old[3].data_ptr(),
                                             float(max_grad), float(model.percent_dense) * float(extent), float(min_opacity),
                                             big_ws, int(N), flags.data_ptr(), sp))
    clone, split = (flags & 1).bool(), (flags & 2).bool()
    prune_self, prune_child = (flags & 4).bool(), (flags & 8).bool()
    idx_keep = torch.nonzero(~split & ~prune_self).view(-1)
    idx_clone = torch.nonzero(clone & ~prune_self).view(-1)
    idx_split = torch.nonzero(split).view(-1)
    S = int(idx_split.numel())
    # the reference's draw (:579-582): stds = get_scaling[selected].repeat(N, 1) with a zero third column
    if _samples is not None:
        samples = _samples.to(dev, torch.float32).contiguous()
    else:
        stds = torch.exp(old[4][idx_split]).repeat(N, 1)
        stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator).contiguous()
    rank = torch.cumsum(split.to(torch.int32), 0) - 1
    idx_child = torch.nonzero(split & ~prune_child).view(-1)
    src = torch.cat([idx_keep, idx_clone] + [idx_child] * N).to(torch.int32).contiguous()
    kind = torch.cat([torch.zeros_like(idx_keep), torch.ones_like(idx_clone)] + [torch.full_like(idx_child, 2)] * N).to(torch.uint8).contiguous()
    sample_row = torch.cat([torch.full_like(idx_keep, -1), torch.full_like(idx_clone, -1)] +
                           [rank[idx_child].to(idx_child.dtype) + c * S for c in range(N)]).to(torch.int32).contiguous()
    P_new = int(src.numel())

    opt = getattr(model, "optimizer", None)
    groups = {g["name"]: g for g in opt.param_groups} if opt is not None else {}
    states = [opt.state.get(groups[name]["params"][0], None) if name in groups else None for name, _ in _GROUPS]
    src_t, dst_t, new = [], [], []
    for (name, attr), o, st in zip(_GROUPS, old, states):
        n = torch.empty((P_new,) + tuple(o.shape[1:]), dtype=torch.float32, device=dev)
        m1 = m2 = nm1 = nm2 = None
        if st is not None:
            m1, m2 = f32c(st["exp_avg"]), f32c(st["exp_avg_sq"])
            nm1, nm2 = torch.empty_like(n), torch.empty_like(n)
        new.append((n, nm1, nm2))
        src_t += [o, m1, m2]
        dst_t += [n, nm1, nm2]
    import ctypes as C
    ptrs = lambda ts: (C.c_void_p * 18)(*[t.data_ptr() if (t is not None and t.numel()) else None for t in ts])
    if P_new > 0:
        with torch.cuda.device(dev):
            _lib.check(_LIB.g4s_densify_gather(P_new, rest_w, src.data_ptr(), kind.data_ptr(), sample_row.data_ptr(),
                                               samples.data_ptr() if samples.numel() else None, int(N), ptrs(src_t), ptrs(dst_t),
                                               torch.cuda.current_stream(dev).cuda_stream))
    # install the new tensors the way _prune_optimizer / cat_tensors_to_optimizer do (:510-527, :545-565)
    for (name, attr), (n, nm1, nm2), st in zip(_GROUPS, new, states):
        param = nn.Parameter(n.requires_grad_(True))
        if name in groups:
            g = groups[name]
            if st is not None:
                del opt.state[g["params"][0]]
                st["exp_avg"], st["exp_avg_sq"] = nm1, nm2
                g["params"][0] = param
                opt.state[param] = st
            else:
                g["params"][0] = param
        setattr(model, attr, param)
    model.xyz_gradient_accum = torch.zeros((P_new, 1), device=dev)
    model.denom = torch.zeros((P_new, 1), device=dev)
    model.max_radii2D = torch.zeros((P_new,), device=dev)
    return P_new
