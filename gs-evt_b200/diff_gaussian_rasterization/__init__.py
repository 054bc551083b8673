"""Drop-in replacement for the reference's `diff_gaussian_rasterization` package
(submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:21-270).

Same public names, argument order, defaults, return tuple and gradient slots:
  GaussianRasterizationSettings  — 18-field NamedTuple (:189-207)
  GaussianRasterizer(nn.Module)  — .forward(...) -> (color, radii, depth, opacity, n_touched), .markVisible (:209-270)
  rasterize_gaussians(...)       — functional form (:21-50)

The work is done by libgsevt.so (hand-written sm_100a kernels) through the C ABI of include/gsevt.h;
PyTorch only owns the memory and the stream.  Differences that do not change results:
  * gradients are produced only for inputs with ctx.needs_input_grad — for GS-EVT's frozen map that
    removes the reference's 352 B/Gaussian of zero-filled gradient tensors (rasterize_points.cu:165-176);
  * dL_dtau / dL_dvel are reduced to the 12 pose floats inside the kernels instead of materialising
    (P,6) tensors and calling torch.sum (:163-169).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from gsevt import lib as _lib

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    angular_vel: torch.Tensor
    linear_vel: torch.Tensor
    vel_transofrm: torch.Tensor
    vel_transofrm_inv: torch.Tensor
    delta_time: float
    debug: bool


def _f32(t, device):
    """Contiguous float32 tensor on `device`, or None for the reference's 'empty tensor' arguments."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


def _fill_common(a, rs, dev, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, keep):
    P = means3D.shape[0]
    a.P = P
    a.sh_degree = int(rs.sh_degree)
    a.sh_coeffs = int(sh.shape[1]) if sh is not None else 0
    a.width = int(rs.image_width)
    a.height = int(rs.image_height)
    a.tanfovx = float(rs.tanfovx)
    a.tanfovy = float(rs.tanfovy)
    a.scale_modifier = float(rs.scale_modifier)
    a.delta_time = float(rs.delta_time)
    a.prefiltered = int(bool(rs.prefiltered))
    a.debug = int(bool(rs.debug))
    for name, t in (("background", rs.bg), ("viewmatrix", rs.viewmatrix), ("projmatrix", rs.projmatrix),
                    ("projmatrix_raw", rs.projmatrix_raw), ("campos", rs.campos),
                    ("vel_transform", rs.vel_transofrm), ("vel_transform_inv", rs.vel_transofrm_inv)):
        t = _f32(t, dev)
        keep.append(t)
        setattr(a, name, _lib.ptr(t))
    for name, t in (("means3D", means3D), ("shs", sh), ("colors_precomp", colors_precomp), ("opacities", opacities),
                    ("scales", scales), ("rotations", rotations), ("cov3D_precomp", cov3Ds_precomp)):
        setattr(a, name, _lib.ptr(t))


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        theta, rho, w, v, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, theta, rho, w, v, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                theta, rho, w, v, raster_settings):
        lib = _lib.load()
        _lib.require_device()
        rs = raster_settings
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:58-60
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("diff_gaussian_rasterization (gsevt): inputs must live on a CUDA device")
        means3D_c = _f32(means3D, dev) if means3D.numel() else means3D
        sh_c, col_c = _f32(sh, dev), _f32(colors_precomp, dev)
        op_c, sc_c, rot_c, cov_c = _f32(opacities, dev), _f32(scales, dev), _f32(rotations, dev), _f32(cov3Ds_precomp, dev)
        P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
        if P != 0:  # with P == 0 every per-Gaussian tensor is empty and the reference returns zeros (rasterize_points.cu:85)
            if (sh_c is None) == (col_c is None):
                raise Exception("Please provide excatly one of either SHs or precomputed colors!")
            if ((sc_c is None or rot_c is None) and cov_c is None) or ((sc_c is not None or rot_c is not None) and cov_c is not None):
                raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")

        with torch.cuda.device(dev):
            f32 = dict(dtype=torch.float32, device=dev)
            color = torch.zeros((3, H, W), **f32) if P == 0 else torch.empty((3, H, W), **f32)
            depth = torch.zeros((1, H, W), **f32) if P == 0 else torch.empty((1, H, W), **f32)
            opacity = torch.zeros((1, H, W), **f32) if P == 0 else torch.empty((1, H, W), **f32)
            radii = torch.zeros((P,), dtype=torch.int32, device=dev)
            n_touched = torch.zeros((P,), dtype=torch.int32, device=dev)
            geomBuffer = torch.empty((0,), dtype=torch.uint8, device=dev)
            binningBuffer = torch.empty((0,), dtype=torch.uint8, device=dev)
            imgBuffer = torch.empty((0,), dtype=torch.uint8, device=dev)
            num_rendered = 0
            if P != 0:
                import ctypes as C
                gb, ib = C.c_size_t(0), C.c_size_t(0)
                _lib.check(lib.gsevt_raster_sizes(P, W, H, C.byref(gb), C.byref(ib)), "gsevt_raster_sizes")
                geomBuffer = torch.empty((gb.value,), dtype=torch.uint8, device=dev)
                imgBuffer = torch.empty((ib.value,), dtype=torch.uint8, device=dev)
                a = _lib.GsevtRasterArgs()
                keep = []
                _fill_common(a, rs, dev, means3D_c, sh_c, col_c, op_c, sc_c, rot_c, cov_c, keep)
                a.want_n_touched = 1
                a.geom_buffer, a.geom_bytes = geomBuffer.data_ptr(), gb.value
                a.img_buffer, a.img_bytes = imgBuffer.data_ptr(), ib.value
                a.out_color, a.out_depth, a.out_opacity = color.data_ptr(), depth.data_ptr(), opacity.data_ptr()
                a.radii, a.n_touched = radii.data_ptr(), n_touched.data_ptr()
                stream = _lib.stream_ptr()
                num_rendered = _lib.check(lib.gsevt_raster_forward_geometry(C.byref(a), stream), "gsevt_raster_forward_geometry")
                bb = lib.gsevt_raster_binning_size(num_rendered)
                binningBuffer = torch.empty((bb,), dtype=torch.uint8, device=dev)
                a.binning_buffer, a.binning_bytes = binningBuffer.data_ptr(), bb
                a.num_rendered = num_rendered
                _lib.check(lib.gsevt_raster_forward_render(C.byref(a), stream), "gsevt_raster_forward_render")

        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.empty = [t is None for t in (col_c, sc_c, rot_c, cov_c, sh_c)]
        ctx.shapes = [None if t is None else tuple(t.shape) for t in (means3D, means2D, sh, colors_precomp, opacities,
                                                                      scales, rotations, cov3Ds_precomp)]
        dummy = torch.empty(0, device=dev)
        ctx.save_for_backward(col_c if col_c is not None else dummy, means3D_c, sc_c if sc_c is not None else dummy,
                              rot_c if rot_c is not None else dummy, cov_c if cov_c is not None else dummy, radii,
                              sh_c if sh_c is not None else dummy, op_c, geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_radii, grad_out_depth, grad_out_opacity, grad_n_touched):
        import ctypes as C
        lib = _lib.load()
        rs = ctx.raster_settings
        (col_c, means3D, sc_c, rot_c, cov_c, radii, sh_c, op_c, geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        col_c, sc_c, rot_c, cov_c, sh_c = [None if e else t for e, t in zip(ctx.empty, (col_c, sc_c, rot_c, cov_c, sh_c))]
        dev = means3D.device
        P = means3D.shape[0]
        need = ctx.needs_input_grad
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            pose = torch.zeros((12,), **f32)
            outs = {}
            if P != 0:
                a = _lib.GsevtRasterArgs()
                keep = []
                _fill_common(a, rs, dev, means3D, sh_c, col_c, op_c, sc_c, rot_c, cov_c, keep)
                gc = grad_out_color if grad_out_color is not None else torch.zeros((3, a.height, a.width), **f32)
                gc = _f32(gc, dev)
                gd = _f32(grad_out_depth, dev) if grad_out_depth is not None else None
                a.dL_dout_color, a.dL_dout_depth = gc.data_ptr(), _lib.ptr(gd)
                a.geom_buffer, a.geom_bytes = geomBuffer.data_ptr(), geomBuffer.numel()
                a.img_buffer, a.img_bytes = imgBuffer.data_ptr(), imgBuffer.numel()
                a.binning_buffer, a.binning_bytes = binningBuffer.data_ptr(), binningBuffer.numel()
                a.num_rendered = ctx.num_rendered
                wb = lib.gsevt_raster_backward_workspace_size(P)
                work = torch.empty((wb,), dtype=torch.uint8, device=dev)
                a.bwd_workspace, a.bwd_workspace_bytes = work.data_ptr(), wb
                a.pose_grads = pose.data_ptr()

                def want(flag, name, shape):
                    if flag:
                        outs[name] = torch.zeros(shape, **f32)
                        setattr(a, name, outs[name].data_ptr())

                M = sh_c.shape[1] if sh_c is not None else 0
                want(need[0], "dL_dmeans3D", (P, 3))
                want(need[1], "dL_dmeans2D", (P, 3))
                want(need[2] and sh_c is not None, "dL_dsh", (P, M, 3))
                want(need[3] and col_c is not None, "dL_dcolors", (P, 3))
                want(need[4], "dL_dopacity", ctx.shapes[4] if ctx.shapes[4] is not None else (P, 1))
                want(need[5] and sc_c is not None, "dL_dscales", (P, 3))
                want(need[6] and rot_c is not None, "dL_drotations", (P, 4))
                want(need[7] and cov_c is not None, "dL_dcov3D", (P, 6))
                _lib.check(lib.gsevt_raster_backward(C.byref(a), _lib.stream_ptr()), "gsevt_raster_backward")

        def g(name):
            return outs.get(name)

        # dgr/diff_gaussian_rasterization/__init__.py:163-169: tau = [rho; theta], vel = [v; w], shaped (1,3)
        grad_rho, grad_theta = pose[0:3].view(1, -1), pose[3:6].view(1, -1)
        grad_v, grad_w = pose[6:9].view(1, -1), pose[9:12].view(1, -1)
        return (g("dL_dmeans3D"), g("dL_dmeans2D"), g("dL_dsh"), g("dL_dcolors"), g("dL_dopacity"), g("dL_dscales"),
                g("dL_drotations"), g("dL_dcov3D"),
                grad_theta if need[8] else None, grad_rho if need[9] else None,
                grad_w if need[10] else None, grad_v if need[11] else None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean (P,) mask of points in front of the near plane (rasterizer_impl.cu:54-66)."""
        with torch.no_grad():
            lib = _lib.load()
            _lib.require_device()
            rs = self.raster_settings
            dev = positions.device
            pos = _f32(positions, dev)
            P = positions.shape[0]
            present = torch.zeros((P,), dtype=torch.bool, device=dev)
            if P != 0:
                vm, pm = _f32(rs.viewmatrix, dev), _f32(rs.projmatrix, dev)
                with torch.cuda.device(dev):
                    _lib.check(lib.gsevt_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                                      present.data_ptr(), _lib.stream_ptr()), "gsevt_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None, w=None, v=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        theta = empty if theta is None else theta
        rho = empty if rho is None else rho
        w = empty if w is None else w
        v = empty if v is None else v

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, w, v, raster_settings)
