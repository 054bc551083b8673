"""Two product paths, one answer: the fused engine against the reference-shaped autograd loop through the drop-in operator.

`engine_vs_operator` evaluates the tracking objective of ONE state (two renders at Exp(-+xi dtau/2) T, normalised
difference against the event frame, 12 pose / velocity gradients) twice:
  * with the native engine (gsevt.engine.TrackingEngine.eval — csrc/api.cu: gsevt_engine_eval), and
  * with RenderFrame -> torch.norm -> loss.backward() through `diff_gaussian_rasterization` (this repo's drop-in operator,
    whose sorted keys / lists / ranges / images / gradients the parity tests pin to the live reference build),
and compares loss, gradients, the two gray images and — bit for bit — the per-tile lists and tile ranges.
bench.py runs it after the timed region so that the benchmarked path at the benchmarked size carries its own evidence
(`parity_check` in the JSON line); tests/test_gpu_parity.py runs it at BASELINE.json's sizes.

The operator's work buffers are read the way the parity tests read them (gsevt_raster_*_offset); no test infrastructure
is involved: both sides are product code.
"""
import numpy as np
import torch

from . import lib as _lib

GRAY = (0.2989, 0.5870, 0.1140)


def rel_max(a, b):
    """max |a - b| relative to the largest component of b."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_comp(a, b, floor=0.05):
    """Per-component relative error with a floor: max_i |a_i - b_i| / max(|b_i|, floor * max|b|).  A small component can
    no longer hide behind the largest one: with floor = 0.05 a component at 5 % of the largest must itself be right to the gate."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-30))
    return float((np.abs(a - b) / den).max())


def _operator_buffers(lib, saved, P, W, H):
    """(sorted keys, sorted ids, ranges) from the drop-in operator's saved work buffers."""
    geom, binning, img = saved[-3], saved[-2], saved[-1]
    base = lambda t: (-t.data_ptr()) % 256
    g = geom.cpu().numpy()
    o = base(geom) + lib.gsevt_raster_geom_offset(b"tiles_touched", P)
    N = int(g[o:o + 4 * P].view(np.uint32).astype(np.int64).sum())
    b = binning.cpu().numpy()
    ok, ol = (base(binning) + lib.gsevt_raster_binning_offset(n, N) for n in (b"point_list_keys", b"point_list"))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    im = img.cpu().numpy()
    orr = base(img) + lib.gsevt_raster_img_offset(b"ranges", W, H)
    return (b[ok:ok + 8 * N].view(np.uint64).copy(), b[ol:ol + 4 * N].view(np.uint32).copy(),
            im[orr:orr + 8 * tiles].view(np.uint32).reshape(tiles, 2).copy())


def view_dict(eng, view, level):
    """The camera block of `view` as the device-side pose algebra of the engine produced it in the most recent evaluation,
    in the argument layout of GaussianRasterizationSettings (dgr/diff_gaussian_rasterization/__init__.py:189-207)."""
    vp = eng.view_params(view)
    raw = np.zeros(16, np.float32)
    raw[0], raw[5], raw[11] = vp["proj_a"], vp["proj_b"], vp["proj_e"]      # the three entries the backward reads (backward.cu:550-560)
    return dict(W=int(eng.width * 0.5 ** level), H=int(eng.height * 0.5 ** level), tanfovx=vp["tanfovx"], tanfovy=vp["tanfovy"],
                viewmatrix=vp["viewmatrix"], projmatrix=vp["projmatrix"], projmatrix_raw=raw, campos=vp["campos"], vel=vp["vel"],
                vel_inv=vp["vel_inv"], delta_time=vp["delta_time"])


def operator_objective(mod, views, act, E, dev, signed=True, background=(0.0, 0.0, 0.0)):
    """The tracking objective through a rasteriser module with the reference's operator API (`mod`: this repo's drop-in
    package, or the reference's own): two rasterisations, gray, normalised difference, norm against E (frame.py:61-94,
    tracker.py:93-103), backward to the 12 gradients [rho, theta, v, w].  act: activated map tensors on `dev`.
    Returns loss, gradients, the two gray images and the saved work buffers of each view."""
    P = act["xyz"].shape[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    pose = {k: torch.zeros(3, device=dev, requires_grad=True) for k in ("theta", "rho", "w", "v")}
    bg = torch.tensor(background, dtype=torch.float32, device=dev)
    zero3 = torch.zeros(3, device=dev)
    grays, saved = [], []
    for v in views:
        rs = mod.GaussianRasterizationSettings(
            image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"], tanfovy=v["tanfovy"], bg=bg, scale_modifier=1.0,
            viewmatrix=t(v["viewmatrix"]).view(4, 4), projmatrix=t(v["projmatrix"]).view(4, 4), projmatrix_raw=t(v["projmatrix_raw"]).view(4, 4),
            sh_degree=3, campos=t(v["campos"]), prefiltered=False, angular_vel=zero3, linear_vel=zero3,
            vel_transofrm=t(v["vel"]).view(4, 4), vel_transofrm_inv=t(v["vel_inv"]).view(4, 4), delta_time=v["delta_time"], debug=False)
        color, radii, depth, opacity, n_touched = mod.GaussianRasterizer(rs)(
            means3D=act["xyz"], means2D=torch.zeros((P, 3), device=dev), opacities=act["opacities"], shs=act["shs"], scales=act["scales"],
            rotations=act["rotations"], theta=pose["theta"], rho=pose["rho"], w=pose["w"], v=pose["v"])
        saved.append(tuple(color.grad_fn.saved_tensors))
        wgt = torch.tensor(GRAY, device=dev).view(3, 1, 1)
        grays.append((color * wgt).sum(dim=0))
    d = grays[1] - grays[0]
    u = d / torch.norm(d, p=2)
    loss = torch.norm(u - E) if signed else torch.norm(torch.abs(u) - torch.abs(E))
    loss.backward()
    g = torch.cat([pose["rho"].grad.view(-1), pose["theta"].grad.view(-1), pose["v"].grad.view(-1), pose["w"].grad.view(-1)]).cpu().numpy()
    return float(loss.detach()), g, [x.detach() for x in grays], saved


def engine_vs_operator(eng, act, state, event_level, level=0, signed=True, lists=True):
    """eng: TrackingEngine with begin_frame() done for the event frame; act: dict of the ACTIVATED map tensors on the engine's
    device (xyz, scales, rotations, opacities, shs); state: (R, T, angular_vel, linear_vel) numpy; event_level: (H, W) tensor
    of pyramid level `level`.  The operator is fed the camera blocks the engine's own pose kernel produced, so the two paths
    rasterise bit-identical inputs: lists, ranges, n_contrib and final_T must then be bit-identical, images and gradients
    agree to rounding.  Returns a dict of error measures."""
    import diff_gaussian_rasterization as ours
    lib = _lib.load()
    dev = eng.device
    R, T, w, v = (np.asarray(x, np.float32) for x in state)
    eng.set_state(R, T, w, v)
    L, g = eng.eval(level, signed)
    gl, gn = eng.gray_images(level)
    Tf, nc = eng.image_state(level)
    views = [view_dict(eng, k, level) for k in (0, 1)]
    bg = (0.0, 0.0, 0.0)
    Lo, ga, grays, saved = operator_objective(ours, views, act, event_level, dev, signed, bg)
    out = {"loss_rel": abs(L - Lo) / max(abs(Lo), 1e-30), "grad_rel_max": rel_max(g, ga), "grad_rel_comp": rel_comp(g, ga),
           "loss": L, "loss_operator_path": Lo,
           "gray_rel_max": max(rel_max(gl.cpu().numpy(), grays[0].cpu().numpy()), rel_max(gn.cpu().numpy(), grays[1].cpu().numpy()))}
    if lists:
        Wl, Hl = views[0]["W"], views[0]["H"]
        same, same_img, n_inst = True, True, 0
        for view in (0, 1):
            keys, ids, ranges = eng.binning(view, level)
            okeys, oids, oranges = _operator_buffers(lib, saved[view], eng.map.P, Wl, Hl)
            same = same and np.array_equal(keys, okeys) and np.array_equal(ids, oids) and np.array_equal(ranges, oranges)
            n_inst += int(keys.size)
            img = saved[view][-1]
            im = img.cpu().numpy()
            base = (-img.data_ptr()) % 256
            oT = im[base + lib.gsevt_raster_img_offset(b"accum_alpha", Wl, Hl):][:4 * Wl * Hl].view(np.uint32)
            on = im[base + lib.gsevt_raster_img_offset(b"n_contrib", Wl, Hl):][:4 * Wl * Hl].view(np.uint32)
            same_img = same_img and np.array_equal(Tf[view].cpu().numpy().view(np.uint32).ravel(), oT) and \
                np.array_equal(nc[view].cpu().numpy().view(np.uint32).ravel(), on)
        out["lists_bit_identical"] = bool(same)
        out["n_contrib_final_T_bit_identical"] = bool(same_img)
        out["instances_compared"] = n_inst
    return out
