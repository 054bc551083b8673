#!/usr/bin/env python3
"""CPU study (plain-C oracle, no GPU) of how well a synthetic scene conditions the tracking objective: correlation between
the normalised rendered intensity change at the TRUE state and the event frame built from events sampled off it, and the
loss along pose offsets (a usable scene has a clear minimum at 0 that is deeper than the event noise).  Test
infrastructure: it imports oracle/, like the tests do.

    python tests/scene_study_cpu.py --gaussians 20000 --structure 300 --fine-shift -3 --events 7500 --model threshold
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT]
from gsevt import synth  # noqa: E402
from oracle import event_oracle as eo  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=20000)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--events", type=int, default=7500)
    ap.add_argument("--structure", type=int, default=0)
    ap.add_argument("--structure-scale", type=float, default=0.25)
    ap.add_argument("--fine-shift", type=float, default=0.0)
    ap.add_argument("--depth", default="3,6")
    ap.add_argument("--model", default="threshold")
    ap.add_argument("--lin-scale", type=float, default=1.0)
    ap.add_argument("--ang-scale", type=float, default=1.0)
    ap.add_argument("--offsets-mm", default="0,1,2,5,10,20,50")
    a = ap.parse_args()
    D = synth.DESK
    W, H = a.width, a.height
    s = W / D["W"]
    fx, fy = D["fx"] * s, D["fy"] * s
    depth = tuple(float(x) for x in a.depth.split(","))
    raw = synth.synth_map(a.gaussians, seed=0, W=W, H=H, fx=fx, fy=fy, structure=a.structure, structure_scale=a.structure_scale,
                          fine_opacity_shift=a.fine_shift, structure_depth=depth)
    act = synth.activate(raw)
    K = np.array([fx, 0, W / 2.0, 0, fy, H / 2.0, 0, 0, 1.0]).reshape(3, 3)
    R, T = np.asarray(D["R"], np.float32).reshape(3, 3), np.asarray(D["T"], np.float32)
    w, v = np.asarray(D["angular_vel"], np.float32) * a.ang_scale, np.asarray(D["linear_vel"], np.float32) * a.lin_scale
    zero = np.zeros((H, W), np.float32)
    _, _, aux = orc.tracking_eval(act, R, T, w, v, 0.05, W, H, fx, fy, 0, zero + 1e-3, True)
    g0, g1 = aux["gray"]
    dI = g1 - g0
    make = synth.threshold_events if a.model == "threshold" else synth.sample_events
    tab = make(dI, a.events, 0, 49999, K, D["dist"], seed=1000)
    E = eo.event_frame(tab[:, 1], tab[:, 2], tab[:, 3], W, H, K, D["dist"])
    E = np.asarray(E[0] if isinstance(E, tuple) else E, np.float32).reshape(H, W)
    u = dI / np.linalg.norm(dI)
    corr = float((u * E).sum())
    out = {"scene": vars(a), "corr_true_state": round(corr, 4), "loss_true_state": round(float(np.linalg.norm(u - E)), 4),
           "gray_mean": round(float(g0.mean()), 4), "dI_rms": float(np.sqrt((dI ** 2).mean())), "flow_px_at_4m": round(float(np.linalg.norm(v) * 0.05 * fx / 4), 3)}
    # loss along a pose offset in camera x and along the velocity direction (events fixed)
    curves = {}
    for name, direction in (("x", np.array([1.0, 0, 0])), ("along_v", v / np.linalg.norm(v))):
        row = []
        for mm in [float(x) for x in a.offsets_mm.split(",")]:
            Lp = []
            for sgn in ((1,) if mm == 0 else (1, -1)):
                To = (T + sgn * direction * mm * 1e-3).astype(np.float32)
                L, _, _ = orc.tracking_eval(act, R, To, w, v, 0.05, W, H, fx, fy, 0, E, True)
                Lp.append(round(float(L), 4))
            row.append((mm, Lp))
        curves[name] = row
    out["loss_vs_offset_mm"] = curves
    print(json.dumps(out))


if __name__ == "__main__":
    main()
