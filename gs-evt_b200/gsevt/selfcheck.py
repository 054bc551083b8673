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


def engine_vs_operator(eng, gaussians, state, delta_tau, event_level0, level=0, signed=True, background=(0.0, 0.0, 0.0), lists=True):
    """eng: TrackingEngine with begin_frame() done for `event_level0`'s frame; gaussians: GaussianModel holding the same map;
    state: (R, T, angular_vel, linear_vel) numpy; event_level0: (1, H, W) tensor of pyramid level `level`.
    Returns a dict of error measures (all relative)."""
    from gaussian_splatting.utils.graphics_utils import focal2fov
    from utils.render_camera.camera import Camera
    from utils.render_camera.frame import RenderFrame
    lib = _lib.load()
    dev = eng.device
    R, T, w, v = (np.asarray(x, np.float32) for x in state)
    eng.set_state(R, T, w, v)
    L, g = eng.eval(level, signed)
    gl, gn = eng.gray_images(level)
    W0, H0 = eng.width, eng.height
    cam = Camera(torch.from_numpy(R.reshape(3, 3).copy()), torch.from_numpy(T.copy()), torch.from_numpy(w.copy()).to(dev),
                 torch.from_numpy(v.copy()).to(dev), focal2fov(eng_fx(eng), W0), focal2fov(eng_fy(eng), H0), W0, H0,
                 delta_tau=delta_tau, device=dev)
    cam.fx, cam.fy = eng_fx(eng), eng_fy(eng)
    for p in (cam.cam_rot_delta, cam.cam_trans_delta, cam.cam_w_delta, cam.cam_v_delta):
        p.requires_grad_(True)
        p.grad = None
    bg = torch.tensor(background, dtype=torch.float32, device=dev)
    rf = RenderFrame(cam, gaussians, None, bg, level)
    E = event_level0
    loss = torch.norm(rf.sign_delta_Ir - E) if signed else torch.norm(rf.unsign_delta_Ir - torch.abs(E))
    saved = [tuple(img.grad_fn.saved_tensors) for img in getattr(rf, "_raw_colors", [])] if lists else []
    loss.backward()
    ga = torch.cat([cam.cam_trans_delta.grad, cam.cam_rot_delta.grad, cam.cam_v_delta.grad, cam.cam_w_delta.grad]).detach().cpu().numpy()
    Lo = float(loss.detach())
    out = {"loss_rel": abs(L - Lo) / max(abs(Lo), 1e-30), "grad_rel_max": rel_max(g, ga), "grad_rel_comp": rel_comp(g, ga),
           "loss": L, "loss_operator_path": Lo}
    grays = getattr(rf, "_grays", None)
    if grays is not None:
        out["gray_rel_max"] = max(rel_max(gl.cpu().numpy(), grays[0].detach().cpu().numpy()),
                                  rel_max(gn.cpu().numpy(), grays[1].detach().cpu().numpy()))
    if lists and saved:
        Wl, Hl = int(W0 * 0.5 ** level), int(H0 * 0.5 ** level)
        same = True
        n_inst = 0
        for view in (0, 1):
            keys, ids, ranges = eng.binning(view, level)
            okeys, oids, oranges = _operator_buffers(lib, saved[view], eng.map.P, Wl, Hl)
            same = same and np.array_equal(keys, okeys) and np.array_equal(ids, oids) and np.array_equal(ranges, oranges)
            n_inst += int(keys.size)
        out["lists_bit_identical"] = bool(same)
        out["instances_compared"] = n_inst
    return out


def eng_fx(eng):
    return float(eng.fx)


def eng_fy(eng):
    return float(eng.fy)
