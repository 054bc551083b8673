#!/usr/bin/env python3
"""Which synthetic scene can a tracker hold?  OUR tracker alone (seconds per run) over a grid of scene generators:
number of large structure splats, opacity of the fine texture on top, event model, trajectory shape.  Prints the
translation / rotation error against the synthetic ground truth every 10th frame and the frame at which the track is
lost (> 10 cm).  Used to pick the scene of the gated long-sequence parity run (DESIGN.md "Sequences").

    python tools/scene_sweep.py --frames 120 [--gaussians 300000] --out gpurun_out/scene_sweep.json
"""
import argparse
import itertools
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
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--gaussians", type=int, default=300000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--events", type=int, default=30000)
    ap.add_argument("--structures", default="0,400,1500")
    ap.add_argument("--fine-shifts", default="0,-3")
    ap.add_argument("--event-models", default="proportional,threshold")
    ap.add_argument("--trajs", default="orbit")
    ap.add_argument("--out", default=None)
    ap.add_argument("--detail", action="store_true", help="per-frame errors, velocity estimate against the truth, iterations")
    ap.add_argument("--no-weighted-velocity", action="store_true", help="diagnostic: skip the velocity mix after frame 5 (NOT the reference's algorithm)")
    ap.add_argument("--structure-scale", type=float, default=0.25)
    ap.add_argument("--vel-lr-scales", default="1")
    ap.add_argument("--depth", default="3,6")
    ap.add_argument("--lin-scale", type=float, default=1.0)
    ap.add_argument("--ang-scale", type=float, default=1.0)
    a = ap.parse_args()
    if a.no_weighted_velocity:
        from gsevt.engine import TrackingEngine
        TrackingEngine.weighted_velocity = lambda self, *args, **kw: None
    import test_gpu_sequence as tgs
    from gsevt import ate
    dev = torch.device("cuda:0")
    rows = []
    grid = itertools.product([int(x) for x in a.structures.split(",")], [float(x) for x in a.fine_shifts.split(",")],
                             a.event_models.split(","), a.trajs.split(","), [float(x) for x in a.vel_lr_scales.split(",")])
    depth = tuple(float(x) for x in a.depth.split(","))
    for structure, shift, model, traj, vlr in grid:
        if structure == 0 and shift != 0:
            continue
        t0 = time.perf_counter()
        raw, table, gt, desc = tgs.make_sequence(dev, a.gaussians, a.width, a.height, a.frames, a.events, ang_scale=a.ang_scale, lin_scale=a.lin_scale,
                                                 structure_scale=a.structure_scale, structure=structure, vel_lr_scale=vlr, structure_depth=depth, fine_opacity_shift=shift, event_model=model, traj=traj)
        with tempfile.TemporaryDirectory() as td:
            ours, iters, opt_s = tgs.run_ours(raw, table, desc, td)
        c = ate.compare(ours, gt)
        tr, ro = np.array(c["trans_per_frame_m"]), np.array(c["rot_per_frame_deg"])
        lost = int(np.argmax(tr > 0.10)) if (tr > 0.10).any() else None
        row = dict(vel_lr_scale=vlr, depth=depth, structure=structure, fine_opacity_shift=shift, event_model=model, traj=traj, frames=int(len(tr)), lost_at_frame=lost,
                   trans_err_mm_every_10th=[round(float(x) * 1e3, 1) for x in tr[::10]],
                   rot_err_deg_every_10th=[round(float(x), 3) for x in ro[::10]],
                   trans_rmse_mm=round(float(np.sqrt((tr ** 2).mean())) * 1e3, 2), ate=ate.ate(ours, gt),
                   iterations_per_level_mean=[round(float(x), 1) for x in iters.mean(0)], optimisation_s=round(opt_s, 2),
                   wall_s=round(time.perf_counter() - t0, 1))
        if a.detail:
            from gsevt import synth
            D = synth.DESK
            gtt = synth.ground_truth_trajectory(a.frames, 0.05, D["R"], D["T"], np.asarray(D["linear_vel"]) * a.lin_scale,
                                                np.asarray(D["angular_vel"]) * a.ang_scale, mode=traj)
            vel = tgs.run_ours.last_tracker.velocities
            ang = lambda x, y: float(np.degrees(np.arccos(np.clip(np.dot(x, y) / (np.linalg.norm(x) * np.linalg.norm(y) + 1e-30), -1, 1))))
            row["per_frame"] = [dict(f=j, trans_mm=round(float(tr[j]) * 1e3, 2), rot_deg=round(float(ro[j]), 4),
                                     v_ratio=round(float(np.linalg.norm(vel[j][1]) / np.linalg.norm(gtt[j][1])), 3),
                                     v_angle_deg=round(ang(vel[j][1], gtt[j][1]), 2),
                                     w_ratio=round(float(np.linalg.norm(vel[j][0]) / np.linalg.norm(gtt[j][2])), 3),
                                     w_angle_deg=round(ang(vel[j][0], gtt[j][2]), 2), iters=[int(x) for x in iters[j]])
                                for j in range(len(tr))]
        print(json.dumps(row), flush=True)
        rows.append(row)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
