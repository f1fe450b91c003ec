"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp32) of GaussianModel.compute_mip_filter,
2d-gaussian-splatting/scene/gaussian_model.py:388-434, line by line.

Only tests/ may import this; the product path (g4splat_b200/gaussian_model.py -> g4s_mip_filter)
never does.  Pinned by tests/golden/mip_filter_*.npz, which tests/golden/make_golden_mip.py produced
by calling the reference's own method on the CPU.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def compute_mip_filter(xyz: np.ndarray, cameras, znear: float = 0.2, filter_variance: float = 0.2) -> np.ndarray:
    xyz = np.asarray(xyz, dtype=f32)
    distance = np.full(xyz.shape[0], 100000.0, dtype=f32)                       # :395
    valid_points = np.zeros(xyz.shape[0], dtype=bool)                            # :396
    focal_length = 0.0
    for cam in cameras:                                                          # :400
        R = np.asarray(cam.R, dtype=f32)
        T = np.asarray(cam.T, dtype=f32)
        xyz_cam = (xyz @ R + T[None, :]).astype(f32)                             # :406
        valid_depth = xyz_cam[:, 2] > f32(znear)                                 # :410
        x, y, z = xyz_cam[:, 0], xyz_cam[:, 1], xyz_cam[:, 2]
        z = np.maximum(z, f32(0.001))                                            # :414
        x = x / z * f32(cam.focal_x) + f32(cam.image_width / 2.0)                # :416
        y = y / z * f32(cam.focal_y) + f32(cam.image_height / 2.0)               # :417
        in_screen = (x >= f32(-0.15 * cam.image_width)) & (x <= f32(cam.image_width * 1.15)) & \
                    (y >= f32(-0.15 * cam.image_height)) & (y <= f32(1.15 * cam.image_height))   # :420
        valid = valid_depth & in_screen                                          # :423
        distance[valid] = np.minimum(distance[valid], z[valid])                  # :426
        valid_points |= valid                                                    # :427
        if focal_length < cam.focal_x:                                           # :428
            focal_length = cam.focal_x
    distance[~valid_points] = distance[valid_points].max()                       # :431
    return (distance / f32(focal_length) * f32(filter_variance ** 0.5))[:, None]  # :433-434
