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


def make_sequence(dev, P, W, H, n_frames, n_events, dtau=0.05, seed=0, ang_scale=5.0, lin_scale=2.0, scale_mult=1.0,
                  structure=0, structure_scale=0.25, fine_opacity_shift=0.0, event_model="proportional", traj="drift",
                  orbit_period=12.0, vel_lr_scale=1.0, structure_depth=(3.0, 6.0)):
    """Synthetic map + ground-truth trajectory + events sampled from the intensity change rendered (by the
    engine) at the true pose / velocity of every frame (SURVEY.md 8(d) "Events").
    structure / fine_opacity_shift: the trackable scene of gsevt.synth.synth_map; event_model "threshold": contrast-threshold
    events (synth.threshold_events) instead of draws proportional to |delta I|; traj "orbit": closed-loop trajectory;
    vel_lr_scale: multiplies the yaml's cam_v_delta / cam_w_delta for BOTH trackers (the normalised loss does not see the
    magnitude of the velocity, so Adam walks along that direction at its full step; DESIGN.md "Sequences")."""
    import torch
    from gsevt import ate, synth
    from gsevt.engine import EventFrameBuilder, PackedMap, TrackingEngine
    D = synth.DESK
    s = W / D["W"]
    fx, fy = D["fx"] * s, D["fy"] * s
    raw = synth.synth_map(P, seed=seed, W=W, H=H, fx=fx, fy=fy, structure=structure, structure_scale=structure_scale,
                          fine_opacity_shift=fine_opacity_shift, structure_depth=structure_depth)
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
    gt = synth.ground_truth_trajectory(n_frames, dtau, D["R"], D["T"], lin, ang, mode=traj, orbit_period=orbit_period)
    make_events = synth.threshold_events if event_model == "threshold" else synth.sample_events
    b = EventFrameBuilder(W, H, K, D["dist"], levels=1, device=dev)
    z = np.zeros(1, np.int16)
    dummy = b.build(z, z, z.astype(np.uint8))
    tabs = []
    for j, (pose, v, w, t) in enumerate(gt):
        eng.set_state(pose[:3, :3].astype(np.float32), pose[:3, 3].astype(np.float32), w.astype(np.float32), v.astype(np.float32))
        eng.begin_frame(dtau, dummy[0], dummy[1])
        eng.eval(0, True)
        gl, gn = eng.gray_images(0)
        tabs.append(make_events((gn - gl).cpu().numpy(), n_events, round(j * dtau * 1e6), round((j + 1) * dtau * 1e6) - 1,
                                K, D["dist"], seed=1000 + j))
    eng.close()
    table = np.concatenate(tabs, 0)
    gt_tum = (np.array([g[3] for g in gt]), np.array([g[0][:3, 3] for g in gt]),
              np.array([ate.matrix_to_quat(g[0][:3, :3]) for g in gt]))
    lr = dict(D["lr"])
    lr["cam_v_delta"], lr["cam_w_delta"] = lr["cam_v_delta"] * vel_lr_scale, lr["cam_w_delta"] * vel_lr_scale
    desc = dict(W=W, H=H, fx=fx, fy=fy, cx=W / 2.0, cy=H / 2.0, dist=list(D["dist"]), R=list(D["R"]), T=list(D["T"]),
                angular_vel=(gt[0][2] if traj == "orbit" else ang).tolist(), linear_vel=(gt[0][1] if traj == "orbit" else lin).tolist(), lr=lr, converged_threshold=D["converged_threshold"],
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
                            Optimizer__cam_v_delta=desc["lr"]["cam_v_delta"], Optimizer__cam_w_delta=desc["lr"]["cam_w_delta"],
                            Tracking__initial_vel={"angular_vel": desc["angular_vel"], "linear_vel": desc["linear_vel"]})
    cpath = os.path.join(work, "config.yaml")
    with open(cpath, "w") as f:
        yaml.safe_dump(cfg, f)
    tr = gs_main.main(cpath)
    run_ours.last_tracker = tr
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


def sequence_report(dev, work, P=40000, W=320, H=240, n_frames=8, n_events=12000, ref_twice=False, **seq_kw):
    from gsevt import ate
    raw, table, gt, desc = make_sequence(dev, P, W, H, n_frames, n_events, **seq_kw)
    ours, it_o, s_o = run_ours(raw, table, desc, work)
    ref, it_r, s_r = run_reference(raw, table, desc, work)
    cmp_ = ate.compare(ours, ref)
    self_spread = None
    if ref_twice:   # the unmodified reference against its own second run on the same files: what "identical" means here
        ref2, it_r2, _ = run_reference(raw, table, desc, os.path.join(work, "second"))
        c2 = ate.compare(ref2, ref)
        self_spread = {"per_frame_trans_m": [round(x, 6) for x in c2["trans_per_frame_m"]],
                       "per_frame_rot_deg": [round(x, 5) for x in c2["rot_per_frame_deg"]], "iterations_second_run": it_r2.tolist()}
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
           "optimisation_seconds": {"ours": round(s_o, 3), "reference": round(s_r, 3)},
           "reference_vs_its_own_second_run": self_spread, "generator": {k: (list(v) if isinstance(v, tuple) else v) for k, v in seq_kw.items()}}
    return rep


@pytest.mark.gpu
def test_sequence_tracking_matches_reference_tracker(built, cuda_dev, tmp_path):
    """Whole pipeline, both implementations, same files in (map.ply, events.txt, config.yaml -> Tracker.tracking()):
    BASELINE.json configs[1] in small — 300 k Gaussians, 640x480, 30 000 events per frame, the desk yaml's own velocities
    and learning rates, 24 event frames (so the velocity mix of tracker.py:246-248 that starts at frame 5 is covered),
    three pyramid levels and both stages per frame with the reference's stopping rule.

    Scene: the TRACKABLE synthetic scene of gsevt.synth (400 large structure splats under a faint fine texture, nothing
    within a metre of the camera plane, contrast-threshold events, closed-loop trajectory).  On the noise-textured map
    used in round 1 both trackers lost the scene after 6-14 frames — a splat crossing the renderer's 0.2 m near cut pops
    over a third of the image and steps the loss (DESIGN.md "Sequences"); on this scene both hold it for as long as they
    are run (profiles/r2_long_sequence_1000f.json) 1-4 mm from the ground truth.

    Gates (measured: profiles/r2_sequence_24f_trackable.json — ours vs reference 0.08-2.07 mm / 0.001-0.027 deg, median
    0.47 mm, 20 of 24 frames under 1 mm; the UNMODIFIED reference against its own second run on the same files 0.01-1.13 mm
    / up to 0.015 deg: the stopping rule thresholds the mean |loss step| on a plateau, so the stop iteration and with it the
    last millimetre jitters with the order of the float atomics in EITHER implementation):
      * rotation: north_star's 0.05 deg on every frame;
      * translation: north_star's 1 mm on the median and on at least 2 frames in 3; no frame further than 3 mm — that is
        2.7x the reference's own repeatability on this scene;
      * ATE against the ground truth (what BASELINE.json's metric calls ATE parity): the two RMSEs within 0.3 mm of each
        other, both trackers within 1 cm of the truth on every frame."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref did not travel with this snapshot")
    rep = sequence_report(cuda_dev, str(tmp_path), P=300000, W=640, H=480, n_frames=24, n_events=30000, ang_scale=1.0, lin_scale=1.0,
                          structure=400, fine_opacity_shift=-3.0, event_model="threshold", traj="orbit")
    print(json.dumps(rep))
    c = rep["ours_vs_reference"]
    assert c["pairs"] == rep["frames"] == 24
    dt, dr = np.array(rep["per_frame_trans_m"]), np.array(rep["per_frame_rot_deg"])
    print("ours vs reference per frame: %s mm, %s deg; optimisation %s s" % ((dt * 1e3).round(2).tolist(), dr.tolist(), rep["optimisation_seconds"]))
    assert dr.max() < 0.05, dr
    assert np.median(dt) < 1e-3 and (dt < 1e-3).sum() >= 16 and dt.max() < 3e-3, dt
    # both trackers actually track: every frame within a centimetre of the synthetic ground truth, equal ATE
    assert max(rep["ours_vs_gt_trans_m"]) < 0.01 and max(rep["reference_vs_gt_trans_m"]) < 0.01
    assert abs(rep["ate_ours"]["ate_rmse_m"] - rep["ate_reference"]["ate_rmse_m"]) < 3e-4, (rep["ate_ours"], rep["ate_reference"])
    it_o, it_r = np.array(rep["iterations_ours"]), np.array(rep["iterations_reference"])
    assert it_o.shape == it_r.shape == (rep["frames"], 3)
    cap = 2 * 200 + 1                                  # coarse + fine stage caps of tracker.py:224-240
    assert it_o.min() >= 1 and it_o.max() <= cap and it_r.max() <= cap
    # same amount of optimisation work per level on average (the stop iteration of a single level jitters, see above)
    assert np.all(np.abs(it_o.mean(0) - it_r.mean(0)) < 0.15 * it_r.mean(0)), (it_o.mean(0), it_r.mean(0))


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
    ap.add_argument("--structure", type=int, default=0)
    ap.add_argument("--fine-shift", type=float, default=0.0)
    ap.add_argument("--event-model", default="proportional")
    ap.add_argument("--traj", default="drift")
    ap.add_argument("--ref-twice", action="store_true")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        rep = sequence_report(torch.device("cuda:0"), td, a.gaussians, a.width, a.height, a.frames, a.events,
                              ang_scale=a.ang_scale, lin_scale=a.lin_scale, scale_mult=a.scale_mult, structure=a.structure,
                              fine_opacity_shift=a.fine_shift, event_model=a.event_model, traj=a.traj, ref_twice=a.ref_twice)
    s = json.dumps(rep)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")
