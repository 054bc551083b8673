#!/usr/bin/env python3
"""Diagnostic: engine forward vs the drop-in operator on identical view parameters — where do the gray images differ?
   python tools/diag_gray.py [P] [W] [H]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT, os.path.join(ROOT, "tests")]
import helpers as H  # noqa: E402
import diff_gaussian_rasterization as ours  # noqa: E402
from gsevt import lib, synth  # noqa: E402
from gsevt.engine import EventFrameBuilder, PackedMap, TrackingEngine  # noqa: E402

P, W, Hh = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 1000000), (2, 640), (3, 480)))
dev = torch.device("cuda:0")
L = lib.load()
sc = H.small_scene(P, W, Hh, seed=1)
A = {k: torch.from_numpy(v).to(dev) for k, v in sc["act"].items()}
eng = TrackingEngine(PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3), W, Hh, sc["fx"], sc["fy"])
eng.set_state(sc["R"], sc["T"], sc["w"], sc["v"])
K = np.array([sc["fx"], 0, W / 2, 0, sc["fy"], Hh / 2, 0, 0, 1.0]).reshape(3, 3)
b = EventFrameBuilder(W, Hh, K, synth.DESK["dist"], device=dev)
z = np.zeros(1, np.int16)
sign, unsign = b.build(z, z, z.astype(np.uint8))
eng.begin_frame(sc["dtau"], sign, unsign)
for mode in (0, 1):
    eng.set_binning(mode)
    eng.eval(0, True)
    gl, gn = eng.gray_images(0)
    T, nc = eng.image_state(0)
    grays = [gl.cpu().numpy(), gn.cpu().numpy()]
    for vi, view in enumerate(sc["views"]):
        vp = eng.view_params(vi)
        print("mode", mode, "view", vi, "view matrix == oracle:", np.array_equal(vp["viewmatrix"], np.asarray(view["viewmatrix"], np.float32).ravel()),
              "proj == oracle:", np.array_equal(vp["projmatrix"], np.asarray(view["projmatrix"], np.float32).ravel()))
        a = H.run_operator(ours, sc, view, dev, bg=(0.0, 0.0, 0.0), want_map_grads=False)
        im = H.parse_our_img(L, a["saved"][-1], W, Hh)
        og = 0.2989 * a["color"][0] + 0.5870 * a["color"][1] + 0.1140 * a["color"][2]
        d = np.abs(grays[vi] - og)
        ncd = nc[vi].cpu().numpy().astype(np.int64) - im["n_contrib"].astype(np.int64)
        Td = T[vi].cpu().numpy() - im["accum_alpha"]
        print("  gray: max abs diff %.3e (max gray %.3f), pixels > 1e-5: %d, > 1e-4: %d" % (d.max(), og.max(), (d > 1e-5).sum(), (d > 1e-4).sum()))
        print("  n_contrib differs at %d pixels (max |diff| %d); final_T bit-identical: %s (max abs diff %.3e)" %
              ((ncd != 0).sum(), np.abs(ncd).max(), np.array_equal(T[vi].cpu().numpy().view(np.uint32), im["accum_alpha"].view(np.uint32)), np.abs(Td).max()))
        ys, xs = np.unravel_index(np.argsort(d.ravel())[::-1][:5], d.shape)
        for y, x in zip(ys, xs):
            print("    pixel (%d, %d) tile (%d, %d): gray %.6f vs %.6f, n_contrib %d vs %d, T %.6e vs %.6e" %
                  (x, y, x // 16, y // 16, grays[vi][y, x], og[y, x], nc[vi][y, x].item(), im["n_contrib"][y, x], T[vi][y, x].item(), im["accum_alpha"][y, x]))
