"""Shared helpers for the parity tests: seeded scenes, settings, parsers for the reference's opaque work
buffers (layout: dgr/cuda_rasterizer/rasterizer_impl.cu:155-194, rasterizer_impl.h:22-28) and for ours
(offsets from the gsevt_raster_*_offset introspection calls)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def desk():
    from gsevt import synth
    return synth.DESK


def small_scene(P, W, H, seed=0, ang_scale=10.0, dtau=0.05, level=0, sh_degree=3):
    """Seeded map (activated) + the two render2 views at the desk pose for a W x H sensor."""
    from gsevt import synth
    from oracle import oracle as orc
    D = synth.DESK
    s = W / D["W"]
    fx, fy = D["fx"] * s, D["fy"] * s
    raw = synth.synth_map(P, seed=seed, W=W, H=H, fx=fx, fy=fy, sh_degree=3)
    act = synth.activate(raw)
    R = np.array(D["R"], np.float32).reshape(3, 3)
    T = np.array(D["T"], np.float32)
    w = np.array(D["angular_vel"], np.float32) * ang_scale
    v = np.array(D["linear_vel"], np.float32)
    views = orc.view_setup(R, T, w, v, dtau, W, H, fx, fy, level)
    return dict(raw=raw, act=act, R=R, T=T, w=w, v=v, views=views, W=W, H=H, fx=fx, fy=fy, dtau=dtau)


def align(x, a=128):
    return (x + a - 1) // a * a


def ref_scan_temp_bytes(P):
    """cub::DeviceScan::InclusiveSum temp size as the reference queries it — obtained from our library,
    which links the same CUB (only used to locate point_offsets, which the tests do not need)."""
    return None


def parse_ref_geom(buf, P):
    """depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb, tiles_touched of the reference's
    geomBuffer (uint8 numpy array)."""
    o, out = 0, {}
    for name, dt, cnt in (("depths", np.float32, P), ("clamped", np.uint8, 3 * P), ("internal_radii", np.int32, P),
                          ("means2D", np.float32, 2 * P), ("cov3D", np.float32, 6 * P), ("conic_opacity", np.float32, 4 * P),
                          ("rgb", np.float32, 3 * P), ("tiles_touched", np.uint32, P)):
        o = align(o)
        n = np.dtype(dt).itemsize * cnt
        out[name] = buf[o:o + n].view(dt).copy()
        o += n
    out["means2D"] = out["means2D"].reshape(P, 2)
    out["cov3D"] = out["cov3D"].reshape(P, 6)
    out["conic_opacity"] = out["conic_opacity"].reshape(P, 4)
    out["rgb"] = out["rgb"].reshape(P, 3)
    out["clamped"] = out["clamped"].reshape(P, 3)
    return out


def parse_ref_binning(buf, N):
    o, out = 0, {}
    for name, dt in (("point_list", np.uint32), ("point_list_unsorted", np.uint32), ("point_list_keys", np.uint64),
                     ("point_list_keys_unsorted", np.uint64)):
        o = align(o)
        n = np.dtype(dt).itemsize * N
        out[name] = buf[o:o + n].view(dt).copy()
        o += n
    return out


def parse_ref_img(buf, W, H):
    HW = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    o, out = 0, {}
    o = align(o); out["accum_alpha"] = buf[o:o + 4 * HW].view(np.float32).reshape(H, W).copy(); o += 4 * HW
    o = align(o); out["n_contrib"] = buf[o:o + 4 * HW].view(np.uint32).reshape(H, W).copy(); o += 4 * HW
    o = align(o); out["ranges"] = buf[o:o + 8 * tiles].view(np.uint32).reshape(tiles, 2).copy()
    return out


def _ours_base(t):
    """Our buffers are carved from the tensor base rounded up to 256 bytes (api.cu base_aligned)."""
    p = t.data_ptr()
    return (p + 255) // 256 * 256 - p


def parse_our_geom(lib, t, P):
    b = t.cpu().numpy()
    base = _ours_base(t)
    off = lambda n: base + lib.gsevt_raster_geom_offset(n.encode(), P)
    rec = b[off("rec"):off("rec") + 32 * P].view(np.float32).reshape(P, 8)
    return dict(rec=rec.copy(),
                rgb4=b[off("rgb4"):off("rgb4") + 16 * P].view(np.float32).reshape(P, 4).copy(),
                cov3D=b[off("cov3D"):off("cov3D") + 24 * P].view(np.float32).reshape(P, 6).copy(),
                radii=b[off("radii"):off("radii") + 4 * P].view(np.int32).copy(),
                clamped=b[off("clamped"):off("clamped") + P].view(np.uint8).copy(),
                tiles_touched=b[off("tiles_touched"):off("tiles_touched") + 4 * P].view(np.uint32).copy(),
                point_offsets=b[off("point_offsets"):off("point_offsets") + 4 * P].view(np.uint32).copy())


def parse_our_binning(lib, t, N):
    b = t.cpu().numpy()
    base = _ours_base(t)
    off = lambda n: base + lib.gsevt_raster_binning_offset(n.encode(), N)
    return dict(point_list_keys=b[off("point_list_keys"):off("point_list_keys") + 8 * N].view(np.uint64).copy(),
                point_list=b[off("point_list"):off("point_list") + 4 * N].view(np.uint32).copy(),
                point_list_keys_unsorted=b[off("point_list_keys_unsorted"):off("point_list_keys_unsorted") + 8 * N].view(np.uint64).copy(),
                point_list_unsorted=b[off("point_list_unsorted"):off("point_list_unsorted") + 4 * N].view(np.uint32).copy())


def parse_our_img(lib, t, W, H):
    b = t.cpu().numpy()
    base = _ours_base(t)
    HW = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    off = lambda n: base + lib.gsevt_raster_img_offset(n.encode(), W, H)
    return dict(accum_alpha=b[off("accum_alpha"):off("accum_alpha") + 4 * HW].view(np.float32).reshape(H, W).copy(),
                n_contrib=b[off("n_contrib"):off("n_contrib") + 4 * HW].view(np.uint32).reshape(H, W).copy(),
                ranges=b[off("ranges"):off("ranges") + 8 * tiles].view(np.uint32).reshape(tiles, 2).copy())


def settings(mod, v, bg, dev, sh_degree=3, w=None, lin=None, debug=False):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    z = np.zeros(3, np.float32)
    return mod.GaussianRasterizationSettings(
        image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"], tanfovy=v["tanfovy"], bg=bg, scale_modifier=1.0,
        viewmatrix=t(v["viewmatrix"]).view(4, 4), projmatrix=t(v["projmatrix"]).view(4, 4),
        projmatrix_raw=t(v["projmatrix_raw"]).view(4, 4), sh_degree=sh_degree, campos=t(v["campos"]), prefiltered=False,
        angular_vel=t(z if w is None else w), linear_vel=t(z if lin is None else lin),
        vel_transofrm=t(v["vel"]).view(4, 4), vel_transofrm_inv=t(v["vel_inv"]).view(4, 4), delta_time=v["delta_time"], debug=debug)


def run_operator(mod, sc, view, dev, bg=(0.1, 0.2, 0.3), dcol=None, ddep=None, sh_degree=3, colors=False, cov_precomp=None,
                 want_map_grads=True):
    """Forward (+ backward when dcol is given) of a rasteriser module (`ours` or the reference) on a scene.
    Returns numpy outputs and, for our operator, the saved work buffers."""
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    act = sc["act"]
    P = act["xyz"].shape[0]
    bgt = t(np.asarray(bg, np.float32))
    leaf = {k: t(act[k]).requires_grad_(want_map_grads) for k in ("xyz", "opacities", "scales", "rotations", "shs")}
    pose = {k: torch.zeros(3, device=dev, requires_grad=True) for k in ("theta", "rho", "w", "v")}
    means2D = torch.zeros((P, 3), device=dev, requires_grad=want_map_grads)
    r = mod.GaussianRasterizer(settings(mod, view, bgt, dev, sh_degree))
    kw = dict(means3D=leaf["xyz"], means2D=means2D, opacities=leaf["opacities"], theta=pose["theta"], rho=pose["rho"],
              w=pose["w"], v=pose["v"])
    col_leaf = None
    if colors:
        col_leaf = t(np.clip(act["shs"][:, 0, :] * 0.28 + 0.5, 0, 1)).requires_grad_(want_map_grads)
        kw["colors_precomp"] = col_leaf
    else:
        kw["shs"] = leaf["shs"]
    if cov_precomp is not None:
        kw["cov3D_precomp"] = t(cov_precomp)
    else:
        kw["scales"], kw["rotations"] = leaf["scales"], leaf["rotations"]
    color, radii, depth, opacity, n_touched = r(**kw)
    saved = tuple(color.grad_fn.saved_tensors)   # keep the work buffers: autograd frees them in backward()
    out = dict(saved=saved, color=color.detach().cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.detach().cpu().numpy(),
               opacity=opacity.detach().cpu().numpy(), n_touched=n_touched.cpu().numpy())
    if dcol is not None:
        loss = (color * t(dcol)).sum()
        if ddep is not None:
            loss = loss + (depth * t(ddep)).sum()
        loss.backward()
        out["pose"] = torch.cat([pose["rho"].grad.view(-1), pose["theta"].grad.view(-1), pose["v"].grad.view(-1),
                                 pose["w"].grad.view(-1)]).cpu().numpy()
        if want_map_grads:
            out["g_means2D"] = means2D.grad.cpu().numpy()
            for k in ("xyz", "opacities") + (() if colors else ("shs",)) + (() if cov_precomp is not None else ("scales", "rotations")):
                out["g_" + k] = leaf[k].grad.cpu().numpy()
            if colors:
                out["g_colors"] = col_leaf.grad.cpu().numpy()
    return out


def rel_max(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype.itemsize == b.dtype.itemsize and np.array_equal(a.view(np.uint8), b.view(np.uint8))
