"""`render()` with the reference's signature and result dictionary
(2d-gaussian-splatting/gaussian_renderer/__init__.py:19-166), built on the B200 rasterizer and the
fused post-processing kernels.  `viewpoint_camera`, `pc` and `pipe` are the reference's own objects
(Camera: scene/cameras.py:18-73, GaussianModel: scene/gaussian_model.py, PipelineParams) -- only the
attributes the reference's render() reads are touched, so they are used duck-typed.

    from g4splat_b200.gaussian_renderer import render      # instead of `from gaussian_renderer import render`
"""
from __future__ import annotations

import math

import torch

from .diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussian_model
from .surface import surface_attributes


def _splat2pix_precomp(viewpoint_camera, pc, scaling_modifier):
    """pipe.compute_cov3D_python: the per-Gaussian 3x3 transform built in torch (:64-75)."""
    W, H = viewpoint_camera.image_width, viewpoint_camera.image_height
    near, far = viewpoint_camera.znear, viewpoint_camera.zfar
    dev = viewpoint_camera.full_proj_transform.device
    ndc2pix = torch.tensor([[W / 2, 0, 0, (W - 1) / 2],
                            [0, H / 2, 0, (H - 1) / 2],
                            [0, 0, far - near, near],
                            [0, 0, 0, 1]], dtype=torch.float32, device=dev).T
    world2pix = viewpoint_camera.full_proj_transform @ ndc2pix
    splat2world = pc.get_covariance(scaling_modifier)
    return (splat2world[:, [0, 1, 3]] @ world2pix[:, [0, 1, 3]]).permute(0, 2, 1).reshape(-1, 9)


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           fused_activations: bool = False):
    """Render the scene; bg_color must be on the GPU.  Same keys as the reference: render,
    viewspace_points, visibility_filter, radii, rend_alpha, rend_normal, rend_normal_cam, rend_dist,
    surf_depth, surf_normal, surf_normal_cam, rend_depth.

    fused_activations (extension, off by default): read the model's raw leaves (_xyz, _features_dc,
    _features_rest, _opacity, _scaling, _rotation, mip_filter) and let the projection kernels apply
    get_scaling / get_rotation / get_opacity / get_features (scene/gaussian_model.py:158-192) in
    registers; gradients arrive at the leaves directly."""
    xyz = pc.get_xyz
    # gradient carrier of the screen-space means (densification statistic), :27-31
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
    )
    rasterizer = GaussianRasterizer(raster_settings=settings)

    if fused_activations and override_color is None and not getattr(pipe, "compute_cov3D_python", False):
        mip = pc.mip_filter if getattr(pc, "use_mip_filter", False) else None
        rendered_image, radii, allmap = rasterize_gaussian_model(
            pc._xyz, screenspace_points, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling, pc._rotation,
            mip, settings)
        rets = {"render": rendered_image, "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0, "radii": radii}
        rets.update(surface_attributes(allmap, viewpoint_camera.world_view_transform,
                                       viewpoint_camera.full_proj_transform, pipe.depth_ratio))
        return rets

    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = _splat2pix_precomp(viewpoint_camera, pc, scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation

    # the reference forces convert_SHs_python off (:79): SH -> RGB always happens in the rasterizer
    shs = colors_precomp = None
    if override_color is None:
        shs = pc.get_features
    else:
        colors_precomp = override_color

    rendered_image, radii, allmap = rasterizer(
        means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc.get_opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)

    rets = {"render": rendered_image, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}
    rets.update(surface_attributes(allmap, viewpoint_camera.world_view_transform,
                                   viewpoint_camera.full_proj_transform, pipe.depth_ratio))
    return rets
