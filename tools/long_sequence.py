#!/usr/bin/env python3
"""BASELINE.json configs[1]-shaped run of OUR tracker alone: a long synthetic event sequence against a 300 k-Gaussian
map at 640x480 through the files and the CLI entry point (map.ply, events.txt, config.yaml -> main.main), like a user of
the reference would run it.  Reports wall time (parsing, per-frame event frames, optimisation), iterations per level,
error against the synthetic ground truth, and peak device memory.  With --reference-frames N the UNMODIFIED
reference tracker runs on the first N frames of the same sequence (it needs ~4 s per frame) and the two trajectories are
compared pose by pose and through their ATE against the ground truth.

    python tools/long_sequence.py --frames 1000 --structure 400 --fine-shift -3 --event-model threshold --traj orbit \
        --ang-scale 1 --lin-scale 1 --reference-frames 200 --out gpurun_out/long_seq.json
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--gaussians", type=int, default=300000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--events", type=int, default=30000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ang-scale", type=float, default=5.0)
    ap.add_argument("--lin-scale", type=float, default=2.0)
    ap.add_argument("--scale-mult", type=float, default=1.0, help="multiplies every Gaussian's extent (smoother texture)")
    ap.add_argument("--structure", type=int, default=0, help="large structure splats of the trackable scene (gsevt.synth.synth_map)")
    ap.add_argument("--fine-shift", type=float, default=0.0)
    ap.add_argument("--event-model", default="proportional")
    ap.add_argument("--traj", default="drift")
    ap.add_argument("--reference-frames", type=int, default=0,
                    help="also run the UNMODIFIED reference tracker (oracle/_ref) on the first N frames of the same files' content")
    a = ap.parse_args()
    import test_gpu_sequence as tgs
    from gsevt import ate
    dev = torch.device("cuda:0")
    t0 = time.perf_counter()
    raw, table, gt, desc = tgs.make_sequence(dev, a.gaussians, a.width, a.height, a.frames, a.events, ang_scale=a.ang_scale,
                                             lin_scale=a.lin_scale, scale_mult=a.scale_mult, structure=a.structure,
                                             fine_opacity_shift=a.fine_shift, event_model=a.event_model, traj=a.traj)
    t_gen = time.perf_counter() - t0
    torch.cuda.reset_peak_memory_stats()
    free0, total = torch.cuda.mem_get_info()
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        ours, iters, opt_s = tgs.run_ours(raw, table, desc, td)
        t_run = time.perf_counter() - t0
    free1, _ = torch.cuda.mem_get_info()
    ref_part = None
    if a.reference_frames > 0:
        n = min(a.reference_frames, a.frames)
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            ref, it_r, s_r = tgs.run_reference(raw, table[:n * a.events], desc, td)
            t_ref = time.perf_counter() - t0
        head = tuple(x[:n] for x in ours)
        gth = tuple(x[:n] for x in gt)
        c = ate.compare(head, ref)
        tr, ro = np.array(c["trans_per_frame_m"]), np.array(c["rot_per_frame_deg"])
        ref_part = {"frames": int(n), "wall_s_reference": round(t_ref, 1), "optimisation_s_reference": round(s_r, 2),
                    "optimisation_s_ours_same_frames": round(float(sum(t for per in tgs.run_ours.last_tracker.iter_counts[:n] for (_, _, _, t) in per)), 2),
                    "ours_vs_reference": {"trans_median_mm": round(float(np.median(tr)) * 1e3, 3), "trans_p90_mm": round(float(np.percentile(tr, 90)) * 1e3, 3),
                                          "trans_max_mm": round(float(tr.max()) * 1e3, 3), "frames_within_1mm": int((tr < 1e-3).sum()),
                                          "rot_median_deg": round(float(np.median(ro)), 5), "rot_max_deg": round(float(ro.max()), 5),
                                          "frames_within_0.05deg": int((ro < 0.05).sum())},
                    "ate_ours_same_frames": ate.ate(head, gth), "ate_reference": ate.ate(ref, gth),
                    "unaligned_vs_gt_rmse_mm": {"ours": round(ate.compare(head, gth)["trans_rmse_m"] * 1e3, 3),
                                                "reference": round(ate.compare(ref, gth)["trans_rmse_m"] * 1e3, 3)},
                    "ours_vs_reference_trans_mm_every_10th": [round(float(x) * 1e3, 2) for x in tr[::10]],
                    "iterations_mean_per_level": {"ours": [round(float(x), 1) for x in iters[:n].mean(0)], "reference": [round(float(x), 1) for x in it_r.mean(0)]}}
    cmp_ = ate.compare(ours, gt)
    rep = {"workload": f"configs[1]-shaped: {a.frames} synthetic event frames x {a.events} events, {a.gaussians}-Gaussian map, {a.width}x{a.height}",
           "frames_tracked": int(len(ours[0])), "wall_s_total": round(t_run, 2), "wall_s_generate_inputs": round(t_gen, 2),
           "optimisation_s": round(opt_s, 3), "frames_per_s_end_to_end": round(len(ours[0]) / t_run, 2),
           "iterations_total": int(iters.sum()), "iterations_per_s_in_optimisation": round(float(iters.sum()) / max(opt_s, 1e-9), 1),
           "iterations_per_level_mean": [round(float(x), 1) for x in iters.mean(0)],
           "iterations_per_level_max": [int(x) for x in iters.max(0)],
           "device_memory_used_MB_after": round((total - free1) / 2**20, 1), "device_memory_used_MB_before": round((total - free0) / 2**20, 1),
           "vs_ground_truth": {k: v for k, v in cmp_.items() if "per_frame" not in k},
           "ate_vs_ground_truth": ate.ate(ours, gt),
           "first_frames_trans_err_m": [round(float(x), 5) for x in cmp_["trans_per_frame_m"][:5]],
           "trans_err_m_every_5th_frame": [round(float(x), 4) for x in cmp_["trans_per_frame_m"][::5]],
           "rot_err_deg_every_5th_frame": [round(float(x), 3) for x in cmp_["rot_per_frame_deg"][::5]],
           "trans_err_mm_percentiles_vs_gt": {str(q): round(float(np.percentile(cmp_["trans_per_frame_m"], q)) * 1e3, 3) for q in (50, 90, 99, 100)},
           "reference_on_first_frames": ref_part,
           "generator": {"ang_scale": a.ang_scale, "lin_scale": a.lin_scale, "scale_mult": a.scale_mult, "structure": a.structure,
                         "fine_opacity_shift": a.fine_shift, "event_model": a.event_model, "traj": a.traj}}
    s = json.dumps(rep)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
