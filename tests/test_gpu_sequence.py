"""Whole-sequence parity (BASELINE.json configs[0]/[1] shape, shortened): this repo's `main.py` (yaml + PLY +
events.txt in, TUM trajectory out, device-side engine) against the UNMODIFIED reference `Tracker.tracking()`
(oracle/_ref, its own CUDA rasteriser + torch autograd, in a subprocess) on the same synthetic sequence with a
known ground-truth trajectory.  Reported: frame-by-frame pose difference (no alignment) and the ATE of both
against ground truth (gsevt.ate).

    python tests/test_gpu_sequence.py [--frames 12 --gaussians 100000 --width 640 --height 480]   # prints the numbers
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gs-evt_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_sequence(dev, P, W, H, n_frames, n_events, dtau=0.05, seed=0, ang_scale=5.0, lin_scale=2.0, scale_mult=1.0):
    """Synthetic map + ground-truth trajectory + events sampled from the intensity change rendered (by the
    engine) at the true pose / velocity of every frame (SURVEY.md 8(d) "Events")."""
    import torch
    from gsevt import ate, synth
    from gsevt.engine import EventFrameBuilder, PackedMap, TrackingEngine
    D = synth.DESK
    s = W / D["W"]
    fx, fy = D["fx"] * s, D["fy"] * s
    raw = synth.synth_map(P, seed=seed, W=W, H=H, fx=fx, fy=fy)
    if scale_mult != 1.0:
        # larger splats = lower-frequency texture: the event frame is blurred 9x9 (event.py:123) while the rendered
        # difference is not, so a map whose texture is finer than the blur carries almost no usable signal
        raw["scaling"] = (raw["scaling"] + np.float32(np.log(scale_mult))).astype(np.float32)
    act = synth.activate(raw)
    A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
    eng = TrackingEngine(PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3), W, H, fx, fy, levels=1)
    K = np.array([fx, 0, W / 2.0, 0, fy, H / 2.0, 0, 0, 1.0]).reshape(3, 3)
    lin = np.asarray(D["linear_vel"]) * lin_scale
    ang = np.asarray(D["angular_vel"]) * ang_scale
    gt = synth.ground_truth_trajectory(n_frames, dtau, D["R"], D["T"], lin, ang)
    b = EventFrameBuilder(W, H, K, D["dist"], levels=1, device=dev)
    z = np.zeros(1, np.int16)
    dummy = b.build(z, z, z.astype(np.uint8))
    tabs = []
    for j, (pose, v, w, t) in enumerate(gt):
        eng.set_state(pose[:3, :3].astype(np.float32), pose[:3, 3].astype(np.float32), w.astype(np.float32), v.astype(np.float32))
        eng.begin_frame(dtau, dummy[0], dummy[1])
        eng.eval(0, True)
        gl, gn = eng.gray_images(0)
        tabs.append(synth.sample_events((gn - gl).cpu().numpy(), n_events, round(j * dtau * 1e6), round((j + 1) * dtau * 1e6) - 1,
                                        K, D["dist"], seed=1000 + j))
    eng.close()
    table = np.concatenate(tabs, 0)
    gt_tum = (np.array([g[3] for g in gt]), np.array([g[0][:3, 3] for g in gt]),
              np.array([ate.matrix_to_quat(g[0][:3, :3]) for g in gt]))
    desc = dict(W=W, H=H, fx=fx, fy=fy, cx=W / 2.0, cy=H / 2.0, dist=list(D["dist"]), R=list(D["R"]), T=list(D["T"]),
                angular_vel=ang.tolist(), linear_vel=lin.tolist(), lr=dict(D["lr"]), converged_threshold=D["converged_threshold"],
                max_optim_iter=D["max_optim_iter"], max_events_per_frame=n_events, background=[0, 0, 0])
    return raw, table, gt_tum, desc


def run_ours(raw, table, desc, work):
    """Through the files and the CLI entry point, exactly like a user of the reference would."""
    import yaml
    from gsevt import ate, synth
    import main as gs_main
    os.makedirs(work, exist_ok=True)
    ply, evs, save = os.path.join(work, "map.ply"), os.path.join(work, "events.txt"), os.path.join(work, "ours")
    synth.save_map_ply(ply, raw)
    synth.write_events_txt(evs, table)
    cfg = synth.make_config(ply, evs, save, W=desc["W"], H=desc["H"], Event__max_events_per_frame=desc["max_events_per_frame"],
                            Tracking__initial_vel={"angular_vel": desc["angular_vel"], "linear_vel": desc["linear_vel"]})
    cpath = os.path.join(work, "config.yaml")
    with open(cpath, "w") as f:
        yaml.safe_dump(cfg, f)
    tr = gs_main.main(cpath)
    tum = ate.load_tum(os.path.join(save, "tracking_pose_tum.txt"))
    iters = np.array([[c + f for (_, c, f, _) in per] for per in tr.iter_counts])
    secs = float(sum(t for per in tr.iter_counts for (_, _, _, t) in per))
    return tum, iters, secs


def run_reference(raw, table, desc, work):
    from oracle import ref_runner
    os.makedirs(work, exist_ok=True)
    inp, out = os.path.join(work, "ref_in.npz"), os.path.join(work, "ref_out.npz")
    d = dict(desc, save_path=os.path.join(work, "ref"))
    np.savez(inp, desc=np.array(d, dtype=object), events=table, **raw)
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "tracker", "--inp", inp, "--out", out],
                   check=True, timeout=3000)
    z = np.load(out)
    t = z["tum"]
    return (t[:, 0], t[:, 1:4], t[:, 4:8]), z["iters"][:, 0].reshape(-1, 3), float(z["opt_time"].sum())


def sequence_report(dev, work, P=40000, W=320, H=240, n_frames=8, n_events=12000, **seq_kw):
    from gsevt import ate
    raw, table, gt, desc = make_sequence(dev, P, W, H, n_frames, n_events, **seq_kw)
    ours, it_o, s_o = run_ours(raw, table, desc, work)
    ref, it_r, s_r = run_reference(raw, table, desc, work)
    cmp_ = ate.compare(ours, ref)
    rep = {"frames": n_frames, "gaussians": P, "size": [W, H], "events_per_frame": n_events,
           "ours_vs_reference": {k: cmp_[k] for k in ("pairs", "trans_max_m", "trans_rmse_m", "rot_max_deg", "rot_mean_deg")},
           "per_frame_trans_m": [round(x, 6) for x in cmp_["trans_per_frame_m"]],
           "per_frame_rot_deg": [round(x, 5) for x in cmp_["rot_per_frame_deg"]],
           "ate_ours": ate.ate(ours, gt), "ate_reference": ate.ate(ref, gt),
           "unaligned_ours_vs_gt": {k: v for k, v in ate.compare(ours, gt).items() if "per_frame" not in k},
           "unaligned_reference_vs_gt": {k: v for k, v in ate.compare(ref, gt).items() if "per_frame" not in k},
           "ours_vs_gt_trans_m": [round(x, 5) for x in ate.compare(ours, gt)["trans_per_frame_m"]],
           "reference_vs_gt_trans_m": [round(x, 5) for x in ate.compare(ref, gt)["trans_per_frame_m"]],
           "iterations_ours": it_o.tolist(), "iterations_reference": it_r.tolist(),
           "optimisation_seconds": {"ours": round(s_o, 3), "reference": round(s_r, 3)}}
    return rep


@pytest.mark.gpu
def test_sequence_tracking_matches_reference_tracker(built, cuda_dev, tmp_path):
    """Whole pipeline, both implementations, same files in (map.ply, events.txt, config.yaml -> Tracker.tracking()):
    BASELINE.json configs[1] in small — 300 k Gaussians, 640x480, 30 000 events per frame, the desk yaml's own
    velocities, 4 event frames, three pyramid levels and both stages per frame with the reference's stopping rule.

    Gate: north_star's 1 mm / 0.05 deg between the two trackers' poses on frames 0-2 (measured over repeated runs:
    0.19-0.34 mm, 0.003-0.008 deg — profiles/r1_sequence_22f_300k_v30.json and the v31 repeats); every frame in the
    same basin (5 cm / 0.5 deg).  Frame 3 is where a level first runs into the 200-iteration cap in BOTH trackers; from
    there the stop iteration jitters with the atomics order (0.9 mm in one run, 3.9 mm in the next), and a few frames
    later both lose the noise-textured synthetic map for good — the reference at frame 6-8, ours at frame 14 in the
    22-frame run — so later frames are reported, not gated.  (On the smaller, faster-rotating 40 k / 320x240 scene this
    test used before, the unmodified reference does not even reproduce its own frame-0 stop iterations:
    profiles/r1_seq_ab_v11.log.)"""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref did not travel with this snapshot")
    rep = sequence_report(cuda_dev, str(tmp_path), P=300000, W=640, H=480, n_frames=4, n_events=30000, ang_scale=1.0, lin_scale=1.0)
    print(json.dumps(rep))
    c = rep["ours_vs_reference"]
    assert c["pairs"] == rep["frames"] == 4
    dt, dr = rep["per_frame_trans_m"], rep["per_frame_rot_deg"]
    print("ours vs reference per frame: %s m, %s deg; optimisation %s s" % (dt, dr, rep["optimisation_seconds"]))
    assert max(dt[:3]) < 1e-3 and max(dr[:3]) < 0.05, (dt, dr)
    assert max(dt) < 0.05 and max(dr) < 0.5, (dt, dr)
    # both trackers actually track these frames (a few centimetres from the synthetic ground truth at most)
    assert max(rep["ours_vs_gt_trans_m"]) < 0.05 and max(rep["reference_vs_gt_trans_m"]) < 0.05
    it_o, it_r = np.array(rep["iterations_ours"]), np.array(rep["iterations_reference"])
    assert it_o.shape == it_r.shape == (rep["frames"], 3)
    cap = 2 * 200 + 1                                  # coarse + fine stage caps of tracker.py:224-240
    assert it_o.min() >= 1 and it_o.max() <= cap and it_r.max() <= cap
    # frame 0: the coarsest level (descending from the same start, far from the plateau) stops at about the same iteration
    assert abs(int(it_o[0][0]) - int(it_r[0][0])) <= 15, (it_o[0], it_r[0])
    for k in ("ate_ours", "ate_reference"):
        assert all(np.isfinite(v) for v in rep[k].values() if isinstance(v, float)), rep[k]


if __name__ == "__main__":
    import argparse
    import tempfile
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--gaussians", type=int, default=40000)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--events", type=int, default=12000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ang-scale", type=float, default=5.0)
    ap.add_argument("--lin-scale", type=float, default=2.0)
    ap.add_argument("--scale-mult", type=float, default=1.0)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        rep = sequence_report(torch.device("cuda:0"), td, a.gaussians, a.width, a.height, a.frames, a.events,
                              ang_scale=a.ang_scale, lin_scale=a.lin_scale, scale_mult=a.scale_mult)
    s = json.dumps(rep)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")
