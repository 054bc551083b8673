#
# Copyright (C) 2023, Inria
# GRAPHDECO research group, https://team.inria.fr/graphdeco
# All rights reserved.
#
# This software is free for non-commercial, research and evaluation use
# under the terms of the LICENSE.md file.
#
# For inquiries contact  george.drettakis@inria.fr
#

import math

import torch
from diff_gaussian_rasterization import (
    GaussianRasterizationSettings,
    GaussianRasterizer,
)

from gaussian_splatting.scene.gaussian_model import GaussianModel
from gaussian_splatting.utils.sh_utils import eval_sh


def render(
    viewpoint_camera,
    pc: GaussianModel,
    pipe,
    bg_color: torch.Tensor,
    scaling_modifier=1.0,
    override_color=None,
    mask=None,
):
    """
    Render the scene.

    Background tensor (bg_color) must be on GPU!
    """

    # Create zero tensor. We will use it to make pytorch return gradients of the 2D (screen-space) means
    if pc.get_xyz.shape[0] == 0:
        return None

    screenspace_points = (
        torch.zeros_like(
            pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda"
        )
        + 0
    )
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    # Set up rasterization configuration
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)

    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
    )

    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity

    # If precomputed 3d covariance is provided, use it. If not, then it will be computed from
    # scaling / rotation by the rasterizer.
    scales = None
    rotations = None
    cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        # check if the covariance is isotropic
        if pc.get_scaling.shape[-1] == 1:
            scales = pc.get_scaling.repeat(1, 3)
        else:
            scales = pc.get_scaling
        rotations = pc.get_rotation

    # If precomputed colors are provided, use them. Otherwise, if it is desired to precompute colors
    # from SHs in Python, do it. If not, then SH -> RGB conversion will be done by rasterizer.
    shs = None
    colors_precomp = None
    if colors_precomp is None:
        if pipe.convert_SHs_python:
            shs_view = pc.get_features.transpose(1, 2).view(
                -1, 3, (pc.max_sh_degree + 1) ** 2
            )
            dir_pp = pc.get_xyz - viewpoint_camera.camera_center.repeat(
                pc.get_features.shape[0], 1
            )
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color

    # Rasterize visible Gaussians to image, obtain their radii (on screen).
    if mask is not None:
        rendered_image, radii, depth, opacity = rasterizer(
            means3D=means3D[mask],
            means2D=means2D[mask],
            shs=shs[mask],
            colors_precomp=colors_precomp[mask] if colors_precomp is not None else None,
            opacities=opacity[mask],
            scales=scales[mask],
            rotations=rotations[mask],
            cov3D_precomp=cov3D_precomp[mask] if cov3D_precomp is not None else None,
            theta=viewpoint_camera.cam_rot_delta,
            rho=viewpoint_camera.cam_trans_delta,
        )
    else:
        rendered_image, radii, depth, opacity, n_touched = rasterizer(
            means3D=means3D,
            means2D=means2D,
            shs=shs,
            colors_precomp=colors_precomp,
            opacities=opacity,
            scales=scales,
            rotations=rotations,
            cov3D_precomp=cov3D_precomp,
            theta=viewpoint_camera.cam_rot_delta,
            rho=viewpoint_camera.cam_trans_delta,
        )

    # Those Gaussians that were frustum culled or had a radius of 0 were not visible.
    # They will be excluded from value updates used in the splitting criteria.
    return {
        "render": rendered_image,
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
        "depth": depth,
        "opacity": opacity,
        "n_touched": n_touched,
    }

def render1(
    curr_viewpoint_camera,
    pc: GaussianModel,
    bg_color: torch.Tensor,
    scaling_modifier=1.0,
    override_color=None,
    mask=None,
):
    """
    Render the scene.

    Background tensor (bg_color) must be on GPU!
    """

    # Create zero tensor. We will use it to make pytorch return gradients of the 2D (screen-space) means
    if pc.get_xyz.shape[0] == 0:
        return None

    screenspace_points = (
        torch.zeros_like(
            pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda"
        )
        + 0
    )
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    # Set up rasterization configuration
    last_vel_transofrm = curr_viewpoint_camera.last_vel_transform.t()
    device = last_vel_transofrm.device
    dtype = last_vel_transofrm.dtype
    T = torch.eye(4, device=device, dtype=dtype)
    R = torch.eye(3, device=device, dtype=dtype)
    t = torch.zeros(3, 1, device=device, dtype=dtype)
    T[:3, :3] = R
    T[:3, 3] = t.squeeze(1)
    curr_vel_transfrm = T
    curr_vel_transfrm_inv = T

    curr_delta_time = 0
    curr_rasterizer = build_rasterizer(curr_viewpoint_camera, curr_vel_transfrm, curr_vel_transfrm_inv,
                                       curr_delta_time, pc, bg_color, scaling_modifier)

    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity

    # Since the 3D gaussian splatting map has been already built up, we directly use the precomputed 3d covariance.
    scales = None
    rotations = None
    cov3D_precomp = None

    # check if the covariance is isotropic
    if pc.get_scaling.shape[-1] == 1:
        scales = pc.get_scaling.repeat(1, 3)
    else:
        scales = pc.get_scaling
    rotations = pc.get_rotation

    # Since the 3D gaussian splatting map has been already built up, we directly use the precomputed colors.
    shs = None
    colors_precomp = None
    if colors_precomp is None:
        shs = pc.get_features
    else:
        colors_precomp = override_color

    theta = curr_viewpoint_camera.cam_rot_delta
    rho = curr_viewpoint_camera.cam_trans_delta
    w = curr_viewpoint_camera.cam_w_delta
    v = curr_viewpoint_camera.cam_v_delta
    curr_render_pkg = run_rasterizer(curr_rasterizer, mask, means3D, means2D, shs, colors_precomp,
                                     opacity, scales, rotations, cov3D_precomp, theta, rho, w, v)

    return curr_render_pkg


def render2(
    last_viewpoint_camera,
    curr_viewpoint_camera,
    next_viewpoint_camera,
    pc: GaussianModel,
    bg_color: torch.Tensor,
    scaling_modifier=1.0,
    override_color=None,
    mask=None,
):
    """
    Render the scene.

    Background tensor (bg_color) must be on GPU!
    """

    # Create zero tensor. We will use it to make pytorch return gradients of the 2D (screen-space) means
    if pc.get_xyz.shape[0] == 0:
        return None

    screenspace_points = (
        torch.zeros_like(
            pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda"
        )
        + 0
    )
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    # Set up rasterization configuration
    last_vel_transofrm = curr_viewpoint_camera.last_vel_transform.t()
    last_vel_transofrm_inv = curr_viewpoint_camera.last_vel_transform_inv.t()
    next_vel_transofrm = curr_viewpoint_camera.next_vel_transform.t()
    next_vel_transofrm_inv = curr_viewpoint_camera.next_vel_transform_inv.t()
    last_delta_time = -curr_viewpoint_camera.delta_tau/ 2
    next_delta_time = curr_viewpoint_camera.delta_tau / 2
    last_rasterizer = build_rasterizer(last_viewpoint_camera, last_vel_transofrm, last_vel_transofrm_inv,
                                       last_delta_time, pc, bg_color, scaling_modifier)
    next_rasterizer = build_rasterizer(next_viewpoint_camera, next_vel_transofrm, next_vel_transofrm_inv,
                                       next_delta_time, pc, bg_color, scaling_modifier)

    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity

    # Since the 3D gaussian splatting map has been already built up, we directly use the precomputed 3d covariance.
    scales = None
    rotations = None
    cov3D_precomp = None

    # check if the covariance is isotropic
    if pc.get_scaling.shape[-1] == 1:
        scales = pc.get_scaling.repeat(1, 3)
    else:
        scales = pc.get_scaling
    rotations = pc.get_rotation

    # Since the 3D gaussian splatting map has been already built up, we directly use the precomputed colors.
    shs = None
    colors_precomp = None
    if colors_precomp is None:
        shs = pc.get_features
    else:
        colors_precomp = override_color

    theta = curr_viewpoint_camera.cam_rot_delta
    rho = curr_viewpoint_camera.cam_trans_delta
    w = curr_viewpoint_camera.cam_w_delta
    v = curr_viewpoint_camera.cam_v_delta
    last_render_pkg = run_rasterizer(last_rasterizer, mask, means3D, means2D, shs, colors_precomp,
                                     opacity, scales, rotations, cov3D_precomp, theta, rho, w, v)
    next_render_pkg = run_rasterizer(next_rasterizer, mask, means3D, means2D, shs, colors_precomp,
                                     opacity, scales, rotations, cov3D_precomp, theta, rho, w, v)

    return last_render_pkg, next_render_pkg


def build_rasterizer(viewpoint_camera, vel_transofrm, vel_transofrm_inv, delta_time, pc, bg_color, scaling_modifier):
    # Set up rasterization configuration
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)

    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        angular_vel=viewpoint_camera.angular_vel,
        linear_vel=viewpoint_camera.linear_vel,
        vel_transofrm=vel_transofrm,
        vel_transofrm_inv=vel_transofrm_inv,
        delta_time=delta_time,
        debug=False,
    )
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    return rasterizer


def run_rasterizer(rasterizer, mask, means3D, means2D, shs, colors_precomp,
                   opacity, scales, rotations, cov3D_precomp, theta, rho, w, v):
    # Rasterize visible Gaussians to image, obtain their radii (on screen).
    n_touched = None
    if mask is not None:
        rendered_image, radii, depth, opacity = rasterizer(
            means3D=means3D[mask],
            means2D=means2D[mask],
            shs=shs[mask],
            colors_precomp=colors_precomp[mask] if colors_precomp is not None else None,
            opacities=opacity[mask],
            scales=scales[mask],
            rotations=rotations[mask],
            cov3D_precomp=cov3D_precomp[mask] if cov3D_precomp is not None else None,
            theta=theta,
            rho=rho,
            w=w,
            v=v,
        )
    else:
        rendered_image, radii, depth, opacity, n_touched = rasterizer(
            means3D=means3D,
            means2D=means2D,
            shs=shs,
            colors_precomp=colors_precomp,
            opacities=opacity,
            scales=scales,
            rotations=rotations,
            cov3D_precomp=cov3D_precomp,
            theta=theta,
            rho=rho,
            w=w,
            v=v,
        )
    # Those Gaussians that were frustum culled or had a radius of 0 were not visible.
    # They will be excluded from value updates used in the splitting criteria.
    return {
        "render": rendered_image,
        "radii": radii,
        "depth": depth,
        "opacity": opacity,
        "n_touched": n_touched,
    }