"""TEST INFRASTRUCTURE -- torch restatement of GaussianModel's parameter activations,
2d-gaussian-splatting/scene/gaussian_model.py:158-192 (get_scaling, get_rotation, get_features,
get_opacity, with and without the mip filter).  Any device / float dtype; gradients by autograd, as in
the reference.

Only tests/ and tests/tools/ may import this; the product path (rasterize_gaussian_model ->
g4s_forward_plan_raw / g4s_backward_raw) never does.  Pinned by tests/golden/activations_*.npz, which
tests/golden/make_golden_activations.py produced by reading the reference GaussianModel's own
properties on the CPU.
"""
from __future__ import annotations

import torch


def get_scaling(_scaling, mip_filter=None):                      # :158-164
    scales = torch.exp(_scaling)
    if mip_filter is not None:
        scales = torch.square(scales) + torch.square(mip_filter)
        scales = torch.sqrt(scales)
    return scales


def get_rotation(_rotation):                                     # :166-168
    return torch.nn.functional.normalize(_rotation)


def get_features(_features_dc, _features_rest):                  # :174-178
    return torch.cat((_features_dc, _features_rest), dim=1)


def get_opacity(_opacity, _scaling, mip_filter=None):            # :180-192
    opacity = torch.sigmoid(_opacity)
    if mip_filter is not None:
        scales = torch.exp(_scaling)
        scales_square = torch.square(scales)
        det1 = scales_square.prod(dim=1)
        scales_after_square = scales_square + torch.square(mip_filter)
        det2 = scales_after_square.prod(dim=1)
        coef = torch.sqrt(det1 / det2)
        opacity = opacity * coef[..., None]
    return opacity


def activate(raw: dict, mip_filter=None) -> dict:
    """raw: _xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation -> the operator's inputs."""
    return dict(means3D=raw["_xyz"], shs=get_features(raw["_features_dc"], raw["_features_rest"]),
                opacities=get_opacity(raw["_opacity"], raw["_scaling"], mip_filter),
                scales=get_scaling(raw["_scaling"], mip_filter), rotations=get_rotation(raw["_rotation"]))
