"""Render glue (mirror of gaussian_splatting/gaussian_renderer/__init__.py:155-384 of the reference):
render1 / render2 / build_rasterizer / run_rasterizer with the same signatures and return dicts.
This is the operator-level (autograd) path; the tracker's hot loop uses gsevt.engine instead."""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def build_rasterizer(viewpoint_camera, vel_transofrm, vel_transofrm_inv, delta_time, pc, bg_color, scaling_modifier):
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
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
    return GaussianRasterizer(raster_settings=settings)


def run_rasterizer(rasterizer, mask, means3D, means2D, shs, colors_precomp, opacity, scales, rotations, cov3D_precomp,
                   theta, rho, w, v):
    sel = (lambda t: t) if mask is None else (lambda t: None if t is None else t[mask])
    rendered_image, radii, depth, opacity_img, n_touched = rasterizer(
        means3D=sel(means3D), means2D=sel(means2D), shs=sel(shs), colors_precomp=sel(colors_precomp),
        opacities=sel(opacity), scales=sel(scales), rotations=sel(rotations), cov3D_precomp=sel(cov3D_precomp),
        theta=theta, rho=rho, w=w, v=v)
    return {"render": rendered_image, "radii": radii, "depth": depth, "opacity": opacity_img,
            "n_touched": n_touched if mask is None else None}


def _map_inputs(pc, override_color):
    scales = pc.get_scaling
    if scales.shape[-1] == 1:
        scales = scales.repeat(1, 3)
    shs, colors = (pc.get_features, None) if override_color is None else (None, override_color)
    return pc.get_xyz, pc.get_opacity, scales, pc.get_rotation, shs, colors


def _screenspace(pc):
    pts = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device=pc.get_xyz.device) + 0
    try:
        pts.retain_grad()
    except Exception:
        pass
    return pts


def render1(curr_viewpoint_camera, pc, bg_color, scaling_modifier=1.0, override_color=None, mask=None):
    """Single view at the current pose (identity velocity transform, delta_time 0)."""
    if pc.get_xyz.shape[0] == 0:
        return None
    eye = torch.eye(4, device=curr_viewpoint_camera.device)
    rasterizer = build_rasterizer(curr_viewpoint_camera, eye, eye, 0, pc, bg_color, scaling_modifier)
    means3D, opacity, scales, rotations, shs, colors = _map_inputs(pc, override_color)
    c = curr_viewpoint_camera
    return run_rasterizer(rasterizer, mask, means3D, _screenspace(pc), shs, colors, opacity, scales, rotations, None,
                          c.cam_rot_delta, c.cam_trans_delta, c.cam_w_delta, c.cam_v_delta)


def render2(last_viewpoint_camera, curr_viewpoint_camera, next_viewpoint_camera, pc, bg_color, scaling_modifier=1.0,
            override_color=None, mask=None):
    """The two half-interval views (t -+ dtau/2) whose difference models the event frame."""
    if pc.get_xyz.shape[0] == 0:
        return None
    c = curr_viewpoint_camera
    last_r = build_rasterizer(last_viewpoint_camera, c.last_vel_transform.t(), c.last_vel_transform_inv.t(),
                              -c.delta_tau / 2, pc, bg_color, scaling_modifier)
    next_r = build_rasterizer(next_viewpoint_camera, c.next_vel_transform.t(), c.next_vel_transform_inv.t(),
                              c.delta_tau / 2, pc, bg_color, scaling_modifier)
    means3D, opacity, scales, rotations, shs, colors = _map_inputs(pc, override_color)
    means2D = _screenspace(pc)
    args = (mask, means3D, means2D, shs, colors, opacity, scales, rotations, None,
            c.cam_rot_delta, c.cam_trans_delta, c.cam_w_delta, c.cam_v_delta)
    return run_rasterizer(last_r, *args), run_rasterizer(next_r, *args)
