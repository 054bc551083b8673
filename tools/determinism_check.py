#!/usr/bin/env python3
"""Bit-determinism of the forward path: the same state evaluated repeatedly (fresh engines too) must give identical
grayscale renders, per-tile lists and tile ranges; the sequence generator must give identical event tables."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT, os.path.join(ROOT, "tests")]


def h(x):
    return hashlib.sha1(np.ascontiguousarray(x).tobytes()).hexdigest()[:12]


def main():
    from gsevt import synth
    from gsevt.engine import EventFrameBuilder, PackedMap, TrackingEngine
    import test_gpu_sequence as tgs
    dev = torch.device("cuda:0")
    for rep in range(2):
        raw, table, gt, desc = tgs.make_sequence(dev, 40000, 320, 240, 3, 12000)
        print("events table", rep, h(table))
    D = synth.DESK
    for (P, W, H) in ((40000, 320, 240), (300000, 640, 480)):
        s = W / D["W"]
        fx, fy = D["fx"] * s, D["fy"] * s
        act = synth.activate(synth.synth_map(P, seed=0, W=W, H=H, fx=fx, fy=fy))
        A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
        K = np.array([fx, 0, W / 2.0, 0, fy, H / 2.0, 0, 0, 1.0]).reshape(3, 3)
        b = EventFrameBuilder(W, H, K, D["dist"], levels=3, device=dev)
        ev = synth.random_events(12000, W, H, 0, 50000, seed=3)
        sign, unsign = b.build(ev[:, 1].astype(np.int16), ev[:, 2].astype(np.int16), ev[:, 3].astype(np.uint8))
        seen = {}
        for fresh in range(3):
            eng = TrackingEngine(PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3), W, H, fx, fy, levels=3)
            eng.set_state(np.array(D["R"], np.float32).reshape(3, 3), np.array(D["T"], np.float32),
                          np.array(D["angular_vel"], np.float32) * 5, np.array(D["linear_vel"], np.float32) * 2)
            eng.begin_frame(0.05, sign, unsign)
            for lvl in (2, 1, 0, 1, 0):
                for again in range(2):
                    L, g = eng.eval(lvl, True)
                    gl, gn = eng.gray_images(lvl)
                    sig = [h(gl.cpu().numpy()), h(gn.cpu().numpy())]
                    for view in (0, 1):
                        keys, lst, rng = eng.binning(view, lvl)
                        sig += [h(keys), h(lst), h(rng)]
                    key = (P, lvl)
                    if key not in seen:
                        seen[key] = sig
                        print("P", P, "level", lvl, "loss", L, sig[:2])
                    elif seen[key] != sig:
                        print("MISMATCH P", P, "level", lvl, "fresh", fresh, "again", again, [a == b_ for a, b_ in zip(seen[key], sig)])
            eng.close()
    print("determinism check done")


if __name__ == "__main__":
    main()
