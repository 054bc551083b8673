#!/usr/bin/env python3
"""BASELINE.json configs[1]-shaped run of OUR tracker alone: a long synthetic event sequence against a 300 k-Gaussian
map at 640x480 through the files and the CLI entry point (map.ply, events.txt, config.yaml -> main.main), like a user of
the reference would run it.  Reports wall time (parsing, per-frame event frames, optimisation), iterations per level,
error against the synthetic ground truth, and peak device memory.  The reference arm is not run here: at ~60 it/s it needs
hours for the same sequence, and pose-by-pose comparison of two trackers on this map is meaningless beyond frame 0
(DESIGN.md section 4); the gated parity lives in tests/.

    python tools/long_sequence.py --frames 300 [--gaussians 300000] [--out gpurun_out/long_seq.json]
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
    a = ap.parse_args()
    import test_gpu_sequence as tgs
    from gsevt import ate
    dev = torch.device("cuda:0")
    t0 = time.perf_counter()
    raw, table, gt, desc = tgs.make_sequence(dev, a.gaussians, a.width, a.height, a.frames, a.events, ang_scale=a.ang_scale,
                                             lin_scale=a.lin_scale, scale_mult=a.scale_mult)
    t_gen = time.perf_counter() - t0
    torch.cuda.reset_peak_memory_stats()
    free0, total = torch.cuda.mem_get_info()
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        ours, iters, opt_s = tgs.run_ours(raw, table, desc, td)
        t_run = time.perf_counter() - t0
    free1, _ = torch.cuda.mem_get_info()
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
           "generator": {"ang_scale": a.ang_scale, "lin_scale": a.lin_scale, "scale_mult": a.scale_mult}}
    s = json.dumps(rep)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
