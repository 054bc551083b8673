#!/usr/bin/env python3
"""Frame-by-frame iteration counts and poses of OUR tracker on the gated sequence test's inputs (no reference run):
a quick A/B harness for kernel variants (e.g. GSEVT_BLEND_BULK=0/1) — identical inputs must give the same stop iterations."""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT, os.path.join(ROOT, "tests")]
import test_gpu_sequence as tgs  # noqa: E402

dev = torch.device("cuda:0")
raw, table, gt, desc = tgs.make_sequence(dev, 300000, 640, 480, 4, 30000, ang_scale=1.0, lin_scale=1.0)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    with tempfile.TemporaryDirectory() as td:
        tum, iters, secs = tgs.run_ours(raw, table, desc, td)
    print(json.dumps({"bulk": os.environ.get("GSEVT_BLEND_BULK", "default"), "iters": iters.tolist(),
                      "T": [[round(float(x), 6) for x in t] for t in tum[1]]}), flush=True)
