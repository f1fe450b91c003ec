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
