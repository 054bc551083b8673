#!/usr/bin/env python3
"""Diagnostic (not a test): runs every CUDA path once on the GPU box and prints mismatch statistics
against the live reference extension (oracle/_ref) and the CPU oracle.  Never asserts."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-evt_b200"))
sys.path.insert(0, ROOT)

from gsevt import synth, lib as glib  # noqa: E402
from gsevt.engine import PackedMap, TrackingEngine, EventFrameBuilder  # noqa: E402
import diff_gaussian_rasterization as ours  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import event_oracle as eo  # noqa: E402

dev = torch.device("cuda:0")
D = synth.DESK


def load_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "ext")
    if not os.path.isdir(p):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dgr", os.path.join(p, "diff_gaussian_rasterization", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(p, "diff_gaussian_rasterization")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_dgr"] = mod
    spec.loader.exec_module(mod)
    return mod


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def settings(mod, v, bg, extra=True):
    return mod.GaussianRasterizationSettings(
        image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"], tanfovy=v["tanfovy"], bg=bg, scale_modifier=1.0,
        viewmatrix=t(v["viewmatrix"]).view(4, 4), projmatrix=t(v["projmatrix"]).view(4, 4),
        projmatrix_raw=t(v["projmatrix_raw"]).view(4, 4), sh_degree=3, campos=t(v["campos"]), prefiltered=False,
        angular_vel=t(np.asarray(D["angular_vel"], np.float32)), linear_vel=t(np.asarray(D["linear_vel"], np.float32)),
        vel_transofrm=t(v["vel"]).view(4, 4), vel_transofrm_inv=t(v["vel_inv"]).view(4, 4), delta_time=v["delta_time"], debug=False)


def cmp(name, a, b, exact=False):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return
    if exact:
        neq = int((a.view(np.uint8) != b.view(np.uint8)).reshape(a.shape[0], -1).any(axis=1).sum()) if a.ndim else int(a != b)
        print(f"  {name}: {'OK bit-exact' if neq == 0 else f'{neq}/{a.shape[0]} rows differ'}")
    else:
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        ref = max(np.abs(b).max(), 1e-30)
        rl2 = np.linalg.norm(d) / max(np.linalg.norm(b.astype(np.float64)), 1e-30)
        print(f"  {name}: max|d|={d.max():.3e} rel_max={d.max() / ref:.3e} rel_l2={rl2:.3e}")


def run_operator(P, W, H, ref):
    print(f"== operator P={P} {W}x{H}")
    m = synth.synth_map(P, seed=0)
    act = synth.activate(m)
    s = W / D["W"]
    R = np.array(D["R"], np.float32).reshape(3, 3)
    T = np.array(D["T"], np.float32)
    views = orc.view_setup(R, T, np.array(D["angular_vel"], np.float32) * 10, np.array(D["linear_vel"], np.float32), 0.05,
                           W, H, D["fx"] * s, D["fy"] * s, 0)
    v = views[1]
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    A = {k: t(x) for k, x in act.items()}
    theta = torch.zeros(3, device=dev, requires_grad=True)
    rho = torch.zeros(3, device=dev, requires_grad=True)
    wv = torch.zeros(3, device=dev, requires_grad=True)
    vv = torch.zeros(3, device=dev, requires_grad=True)
    gen = torch.Generator(device="cpu").manual_seed(5)
    dcol = torch.randn((3, H, W), generator=gen).to(dev)
    ddep = torch.randn((1, H, W), generator=gen).to(dev) * 0.1

    def run(mod, label):
        for x in (theta, rho, wv, vv):
            x.grad = None
        means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
        leaf = {k: A[k].clone().requires_grad_(True) for k in ("xyz", "opacities", "scales", "rotations", "shs")}
        r = mod.GaussianRasterizer(settings(mod, v, bg))
        torch.cuda.synchronize()
        t0 = time.time()
        color, radii, depth, opacity, n_touched = r(means3D=leaf["xyz"], means2D=means2D, opacities=leaf["opacities"], shs=leaf["shs"],
                                                    scales=leaf["scales"], rotations=leaf["rotations"], theta=theta, rho=rho, w=wv, v=vv)
        torch.cuda.synchronize()
        t1 = time.time()
        loss = (color * dcol).sum() + (depth * ddep).sum()
        loss.backward()
        torch.cuda.synchronize()
        t2 = time.time()
        print(f"  [{label}] fwd {1e3 * (t1 - t0):.2f} ms  bwd {1e3 * (t2 - t1):.2f} ms")
        out = dict(color=color, radii=radii, depth=depth, opacity=opacity, n_touched=n_touched,
                   pose=torch.cat([rho.grad.view(-1), theta.grad.view(-1), vv.grad.view(-1), wv.grad.view(-1)]),
                   means2D=means2D.grad, **{"g_" + k: leaf[k].grad for k in leaf})
        return {k: (x.detach().cpu().numpy() if x is not None else None) for k, x in out.items()}

    o = run(ours, "ours")
    run(ours, "ours warm")
    # CPU oracle
    sc = orc.Scene(W, H, v["tanfovx"], v["tanfovy"], bg.cpu().numpy(), act["xyz"], act["opacities"], v["viewmatrix"], v["projmatrix"],
                   v["campos"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"], projmatrix_raw=v["projmatrix_raw"],
                   vel=v["vel"], vel_inv=v["vel_inv"], delta_time=v["delta_time"])
    t0 = time.time()
    fw = orc.forward(sc)
    bw = orc.backward(sc, fw, dcol.cpu().numpy(), ddep.cpu().numpy())
    print(f"  [oracle] fwd+bwd {time.time() - t0:.2f} s, N={fw['num_rendered']}")
    print(" ours vs CPU oracle:")
    cmp("radii", o["radii"], fw["radii"], exact=True)
    cmp("color", o["color"], fw["color"])
    cmp("depth", o["depth"], fw["depth"])
    cmp("opacity", o["opacity"], fw["opacity"])
    cmp("n_touched", o["n_touched"], fw["n_touched"], exact=True)
    cmp("pose_grads", o["pose"], bw["pose_grads"])
    print("   ours pose", o["pose"])
    print("   orc  pose", bw["pose_grads"])
    cmp("dL_dmeans3D", o["g_xyz"], bw["dL_dmeans3D"])
    cmp("dL_dmeans2D", o["means2D"][:, :2], bw["dL_dmean2D"])
    cmp("dL_dopacity", o["g_opacities"].reshape(-1), bw["dL_dopacity"])
    if ref is not None:
        r = run(ref, "reference")
        run(ref, "reference warm")
        print(" ours vs live reference:")
        for k in ("radii", "n_touched"):
            cmp(k, o[k], r[k], exact=True)
        for k in ("color", "depth", "opacity"):
            cmp(k, o[k], r[k])
            cmp(k + " (bits)", o[k].reshape(-1, 1), r[k].reshape(-1, 1), exact=True)
        cmp("pose_grads", o["pose"], r["pose"])
        print("   ref  pose", r["pose"])
        for k in ("means2D", "g_xyz", "g_opacities", "g_scales", "g_rotations", "g_shs"):
            cmp(k, o[k], r[k])
        # buffers: keys / point_list / ranges bit-exact
        try:
            compare_buffers(ref, v, bg, A, P, W, H)
        except Exception:
            traceback.print_exc()


def compare_buffers(ref, v, bg, A, P, W, H):
    import ctypes as C
    lib = glib.load()
    rs = settings(ref, v, bg)
    args = (rs.bg, A["xyz"], torch.Tensor([]), A["opacities"], A["scales"], A["rotations"], rs.scale_modifier, torch.Tensor([]),
            rs.viewmatrix, rs.projmatrix, rs.projmatrix_raw, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, A["shs"],
            rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
    n, color, radii, geomB, binB, imgB, depth, opac, nt = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()

    def obtain(base, off, count, dtype, align=128):
        isz = np.dtype(dtype).itemsize
        start = (base + off + align - 1) // align * align - base
        return start, start + count * isz

    def carve(buf, spec):
        base = buf.data_ptr()
        off, out = 0, {}
        raw = buf.cpu().numpy()
        for name, count, dtype in spec:
            if count is None:
                break
            s0, s1 = obtain(base, off, count, dtype)
            out[name] = raw[s0:s1].view(dtype)
            off = s1
        return out

    g = carve(geomB, [("depths", P, np.float32), ("clamped", 3 * P, np.uint8), ("radii", P, np.int32), ("means2D", 2 * P, np.float32),
                      ("cov3D", 6 * P, np.float32), ("conic_opacity", 4 * P, np.float32), ("rgb", 3 * P, np.float32),
                      ("tiles_touched", P, np.uint32)])
    b = carve(binB, [("point_list", n, np.uint32), ("point_list_unsorted", n, np.uint32), ("keys", n, np.uint64), ("keys_unsorted", n, np.uint64)])
    im = carve(imgB, [("accum_alpha", W * H, np.float32), ("n_contrib", W * H, np.uint32), ("ranges", 2 * W * H, np.uint32)])
    # ours, through the C ABI directly
    a = glib.GsevtRasterArgs()
    keep = []
    ours._fill_common(a, settings(ours, v, bg), dev, A["xyz"], A["shs"], None, A["opacities"], A["scales"], A["rotations"], None, keep)
    gb, ib = C.c_size_t(0), C.c_size_t(0)
    lib.gsevt_raster_sizes(P, W, H, C.byref(gb), C.byref(ib))
    G = torch.zeros(gb.value, dtype=torch.uint8, device=dev)
    I = torch.zeros(ib.value, dtype=torch.uint8, device=dev)
    col = torch.empty((3, H, W), device=dev); dep = torch.empty((1, H, W), device=dev); opa = torch.empty((1, H, W), device=dev)
    rad = torch.zeros(P, dtype=torch.int32, device=dev)
    a.geom_buffer, a.geom_bytes, a.img_buffer, a.img_bytes = G.data_ptr(), gb.value, I.data_ptr(), ib.value
    a.out_color, a.out_depth, a.out_opacity, a.radii = col.data_ptr(), dep.data_ptr(), opa.data_ptr(), rad.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    n2 = lib.gsevt_raster_forward_geometry(C.byref(a), st)
    bb = lib.gsevt_raster_binning_size(n2)
    B = torch.zeros(bb, dtype=torch.uint8, device=dev)
    a.binning_buffer, a.binning_bytes, a.num_rendered = B.data_ptr(), bb, n2
    rc = lib.gsevt_raster_forward_render(C.byref(a), st)
    torch.cuda.synchronize()
    print(f" buffers: num_rendered ref={n} ours={n2} rc={rc}")

    def mine(buf, name, count, dtype, offfn, *extra):
        off = offfn(name.encode(), *extra)
        base = buf.data_ptr()
        al = (base + 255) // 256 * 256 - base
        raw = buf.cpu().numpy()
        return raw[al + off: al + off + count * np.dtype(dtype).itemsize].view(dtype)

    rec = mine(G, "rec", 8 * P, np.float32, lib.gsevt_raster_geom_offset, P).reshape(P, 8)
    rgb4 = mine(G, "rgb4", 4 * P, np.float32, lib.gsevt_raster_geom_offset, P).reshape(P, 4)
    cov = mine(G, "cov3D", 6 * P, np.float32, lib.gsevt_raster_geom_offset, P).reshape(P, 6)
    radm = mine(G, "radii", P, np.int32, lib.gsevt_raster_geom_offset, P)
    tt = mine(G, "tiles_touched", P, np.uint32, lib.gsevt_raster_geom_offset, P)
    vis = g["radii"] > 0
    cmp("radii", radm, g["radii"], exact=True)
    cmp("tiles_touched", tt, g["tiles_touched"], exact=True)
    cmp("depths[vis]", rec[vis, 7], g["depths"][vis], exact=True)
    cmp("means2D[vis]", rec[vis, 0:2], g["means2D"].reshape(P, 2)[vis], exact=True)
    co = g["conic_opacity"].reshape(P, 4)
    cmp("conic[vis]", np.ascontiguousarray(rec[vis][:, [2, 3, 4]]), np.ascontiguousarray(co[vis][:, :3]), exact=True)
    cmp("opacity[vis]", rec[vis, 5], co[vis, 3], exact=True)
    dvis = g["depths"] != 0
    cmp("cov3D[vis]", cov[vis], g["cov3D"].reshape(P, 6)[vis], exact=True)
    cmp("rgb[vis] bits", np.ascontiguousarray(rgb4[vis][:, :3]), g["rgb"].reshape(P, 3)[vis], exact=True)
    cmp("rgb[vis]", rgb4[vis][:, :3], g["rgb"].reshape(P, 3)[vis])
    if n == n2 and n > 0:
        keys = mine(B, "point_list_keys", n, np.uint64, lib.gsevt_raster_binning_offset, n)
        pl = mine(B, "point_list", n, np.uint32, lib.gsevt_raster_binning_offset, n)
        cmp("sorted keys", keys, b["keys"], exact=True)
        cmp("point_list", pl, b["point_list"], exact=True)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        rng = mine(I, "ranges", 2 * tiles, np.uint32, lib.gsevt_raster_img_offset, W, H)
        cmp("ranges", rng, im["ranges"][:2 * tiles], exact=True)
        fT = mine(I, "accum_alpha", W * H, np.float32, lib.gsevt_raster_img_offset, W, H)
        nc = mine(I, "n_contrib", W * H, np.uint32, lib.gsevt_raster_img_offset, W, H)
        cmp("final_T", fT, im["accum_alpha"], exact=True)
        cmp("n_contrib", nc, im["n_contrib"], exact=True)


def run_events():
    print("== events")
    W, H = 640, 480
    K = np.array([D["fx"], 0, D["cx"], 0, D["fy"], D["cy"], 0, 0, 1.0]).reshape(3, 3)
    ev = synth.random_events(30000, W, H, 0, 50000, seed=3)
    b = EventFrameBuilder(W, H, K, D["dist"], levels=3, device=dev)
    x, y, p = ev[:, 1].astype(np.int16), ev[:, 2].astype(np.int16), ev[:, 3].astype(np.uint8)
    torch.cuda.synchronize()
    t0 = time.time()
    sign, unsign = b.build(x, y, p)
    torch.cuda.synchronize()
    print(f"  gpu event frame {1e3 * (time.time() - t0):.2f} ms (first call)")
    t0 = time.time()
    sign, unsign = b.build(x, y, p)
    torch.cuda.synchronize()
    print(f"  gpu event frame {1e3 * (time.time() - t0):.2f} ms (incl. H2D of events)")
    cnt = eo.accumulate(x, y, p, W, H)
    cmp("counts", b.counts.cpu().numpy(), cnt, exact=True)
    s_ref, u_ref = eo.event_frame(x, y, p, W, H, K, D["dist"])
    for l, (sr, ur) in enumerate(zip(eo.pyramid(s_ref[0]), eo.pyramid(u_ref[0]))):
        cmp(f"sign L{l}", b.level_view(sign, l)[0].cpu().numpy(), sr, exact=True)
        cmp(f"unsign L{l}", b.level_view(unsign, l)[0].cpu().numpy(), ur, exact=True)
    try:
        import cv2
        f = np.zeros((H, W), np.float32)
        np.add.at(f, (y.astype(int), x.astype(int)), np.where(p != 0, 1, -1))
        c = cv2.normalize(cv2.GaussianBlur(cv2.undistort(f, K, np.array(D["dist"])), (9, 9), 0, borderType=cv2.BORDER_REPLICATE), None)
        cmp("sign L0 vs cv2", b.level_view(sign, 0)[0].cpu().numpy(), c, exact=True)
    except Exception as e:
        print("  cv2 compare skipped:", e)


def run_engine(P, ref_too=True):
    print(f"== engine P={P}")
    W, H = 640, 480
    m = synth.synth_map(P, seed=0)
    act = synth.activate(m)
    A = {k: t(x) for k, x in act.items()}
    pm = PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3)
    eng = TrackingEngine(pm, W, H, D["fx"], D["fy"])
    R = np.array(D["R"], np.float32).reshape(3, 3)
    T = np.array(D["T"], np.float32)
    w = np.array(D["angular_vel"], np.float32) * 10
    v = np.array(D["linear_vel"], np.float32)
    eng.set_state(R, T, w, v)
    K = np.array([D["fx"], 0, W / 2, 0, D["fy"], H / 2, 0, 0, 1.0]).reshape(3, 3)
    b = EventFrameBuilder(W, H, K, D["dist"], levels=3, device=dev)
    ev = synth.random_events(30000, W, H, 0, 50000, seed=3)
    sign, unsign = b.build(ev[:, 1].astype(np.int16), ev[:, 2].astype(np.int16), ev[:, 3].astype(np.uint8))
    eng.begin_frame(0.05, sign, unsign)
    for level in (2, 0) if P > 50000 else (2, 1, 0):
        for signed in (True, False):
            L, g = eng.eval(level, signed)
            st = eng.status()
            print(f"  eval L{level} signed={signed}: loss={L:.6f} N={list(st.num_rendered)} grads={np.array2string(g, precision=5)}")
            if P <= 50000:
                E = b.level_view(sign, level)[0].cpu().numpy()
                Lo, go, aux = orc.tracking_eval(act, R, T, w, v, 0.05, W, H, D["fx"], D["fy"], level, E, signed)
                print(f"     oracle        : loss={Lo:.6f} grads={np.array2string(go, precision=5)}")
                cmp("     grads vs oracle", g, go)
                gl, gn = eng.gray_images(level)
                cmp("     gray_last", gl.cpu().numpy(), aux["gray"][0])
                cmp("     gray_next", gn.cpu().numpy(), aux["gray"][1])
    # timing of full iterations
    eng.set_state(R, T, w, v)
    eng.begin_frame(0.05, sign, unsign)
    eng.begin_level(0, True)
    eng.iterate(5)
    eng.stream.synchronize()
    st = eng.status()
    print(f"  after 5 iters: done={st.level_done} optim_iter={st.optim_iter} loss={st.last_loss:.6f} N={list(st.num_rendered)}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.begin_level(0, True)
    with torch.cuda.stream(eng.stream):
        e0.record(eng.stream)
        eng.iterate(50)
        e1.record(eng.stream)
    eng.stream.synchronize()
    st = eng.status()
    print(f"  50 iterations: {e0.elapsed_time(e1) / 50:.3f} ms/iter  (executed {st.iters_executed}, done={st.level_done})")
    print("  losses:", np.array2string(eng.losses()[:12], precision=5))
    print("  state:", [np.array2string(x, precision=5) for x in eng.get_state()[1:]])
    t0 = time.time()
    st = eng.run_level(2, False)
    print(f"  run_level(2): optim_iter={st.optim_iter} start_vel={st.start_vel_opt_iter} executed={st.iters_executed} in {time.time() - t0:.3f}s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "arch", glib.load().gsevt_device_arch())
    ref = load_ref()
    print("reference extension:", "loaded" if ref is not None else "absent")
    steps = [lambda: run_events(), lambda: run_operator(20000, 320, 240, ref), lambda: run_engine(20000),
             lambda: run_operator(300000, 640, 480, ref), lambda: run_engine(1000000)]
    for fn in steps:
        try:
            fn()
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()
