#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF (run on a GPU box where
oracle/_ref was built):   python tests/golden/make_golden.py [outdir]

  raster_2k_64x48.npz      the unmodified reference extension on a seeded 2 000-Gaussian scene, both
                           render2 views: radii, sorted keys, point list, tile ranges, n_contrib, final_T,
                           images, the 12 pose/velocity gradients and per-Gaussian gradients for seeded
                           upstream gradients.
  track_eval_4k_160x120.npz   the unmodified reference PIPELINE (Camera / RenderFrame / tracking_loss /
                           autograd, in its own process) on a seeded 4 000-Gaussian scene: the event frame
                           built by its numpy + OpenCV code, loss and the 12 gradients for the signed and
                           unsigned objective at pyramid levels 0 and 1, and 8 optimiser iterations.
The inputs are regenerated from seeds by tests/helpers.py (small_scene) so only outputs are stored.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gs-evt_b200"))
sys.path.insert(0, ROOT)


def raster_golden(out):
    import torch
    import helpers as H
    from conftest import load_reference_extension
    ref = load_reference_extension()
    assert ref is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    P, W, Hh, seed = 2000, 64, 48, 21
    sc = H.small_scene(P, W, Hh, seed=seed)
    rng = np.random.default_rng(17)
    dcol = rng.normal(size=(3, Hh, W)).astype(np.float32)
    ddep = (0.1 * rng.normal(size=(1, Hh, W))).astype(np.float32)
    g = dict(P=P, W=W, H=Hh, seed=seed, dcol=dcol, ddep=ddep)
    for i, view in enumerate(sc["views"]):
        b = H.run_operator(ref, sc, view, dev, dcol=dcol, ddep=ddep)
        sb = b["saved"]
        rg = H.parse_ref_geom(sb[-3].cpu().numpy(), P)
        N = int(rg["tiles_touched"].sum())
        rb = H.parse_ref_binning(sb[-2].cpu().numpy(), N)
        ri = H.parse_ref_img(sb[-1].cpu().numpy(), W, Hh)
        vis = b["radii"] > 0
        for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
            rg[k][~vis] = 0  # culled entries are uninitialised memory in the reference's buffer
        g.update({f"v{i}_num_rendered": N, f"v{i}_radii": b["radii"], f"v{i}_keys": rb["point_list_keys"],
                  f"v{i}_point_list": rb["point_list"], f"v{i}_ranges": ri["ranges"], f"v{i}_n_contrib": ri["n_contrib"],
                  f"v{i}_final_T": ri["accum_alpha"], f"v{i}_color": b["color"], f"v{i}_depth": b["depth"],
                  f"v{i}_opacity": b["opacity"], f"v{i}_n_touched": b["n_touched"], f"v{i}_pose": b["pose"],
                  f"v{i}_depths": rg["depths"], f"v{i}_means2D": rg["means2D"], f"v{i}_conic_opacity": rg["conic_opacity"],
                  f"v{i}_rgb": rg["rgb"], f"v{i}_tiles_touched": rg["tiles_touched"],
                  f"v{i}_g_xyz": b["g_xyz"], f"v{i}_g_means2D": b["g_means2D"], f"v{i}_g_opacities": b["g_opacities"]})
    np.savez_compressed(os.path.join(out, "raster_2k_64x48.npz"), **g)
    print("raster golden: N =", [int(g[f"v{i}_num_rendered"]) for i in range(2)])


def track_golden(out):
    import helpers as H
    from gsevt import synth
    P, W, Hh, seed = 4000, 160, 120, 22
    sc = H.small_scene(P, W, Hh, seed=seed)
    D = synth.DESK
    ev = synth.random_events(4000, W, Hh, 0, 50000, seed=23)
    base = dict(W=W, H=Hh, fx=sc["fx"], fy=sc["fy"], cx=W / 2.0, cy=Hh / 2.0, dist=list(D["dist"]), R=sc["R"].ravel().tolist(),
                T=sc["T"].tolist(), angular_vel=sc["w"].tolist(), linear_vel=sc["v"].tolist(), lr=dict(D["lr"]),
                max_events_per_frame=4000)
    res = dict(P=P, W=W, H=Hh, seed=seed, ev_seed=23, n_events=4000)
    runs = {"eval": dict(plan=[(0, 1, 1), (0, 0, 1), (1, 1, 1), (1, 0, 1)], step=False),
            "iter": dict(plan=[(1, 0, 4), (0, 1, 8)], step=True)}
    for name, extra in runs.items():
        inp, outp = f"/tmp/golden_{name}_in.npz", f"/tmp/golden_{name}_out.npz"
        np.savez(inp, desc=np.array(dict(base, **extra), dtype=object), events=ev, **sc["raw"])
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "iterations", "--inp", inp, "--out", outp], check=True)
        z = np.load(outp)
        for k in z.files:
            res[f"{name}_{k}"] = z[k]
    np.savez_compressed(os.path.join(out, "track_eval_4k_160x120.npz"), **res)
    print("track golden: eval losses", {k: res[k] for k in res if k.startswith("eval_loss")})


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else HERE
    os.makedirs(out, exist_ok=True)
    raster_golden(out)
    track_golden(out)
