"""Diagnostic: n engines on one GPU as tile-split ranks; prints per-rank results and timings."""
import os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gs-evt_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import test_gpu_tilesplit as T
from gsevt import tilesplit

dev = torch.device("cuda:0")
for n in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 8]:
    sc, pm, engs = T._scene_engines(dev, n + 1)
    whole, ranks = engs[0], engs[1:]
    L0, g0 = whole.eval(0, True)
    tilesplit.attach_local(ranks, timeout_s=3.0)
    t0 = time.time()
    stamps = {}
    def fn(r, e):
        a = time.time() - t0
        out = e.eval(0, True)
        stamps[r] = (a, time.time() - t0)
        return out
    res = T._collective(ranks, fn)
    print(f"n={n} unsplit L={L0:.7f}")
    for r, e in enumerate(ranks):
        print("  rank", r, "L=%.7f" % res[r][0], e.split_info(), "t=(%.3f, %.3f)" % stamps[r], "poll", e.poll_done(), flush=True)
    for e in engs:
        e.close()
