#!/usr/bin/env python3
"""Generates tests/golden/events_96x64.npz by running the REFERENCE's own utils/event_camera/event.py
(load_events_from_txt + EventFrame on the CPU: Python scatter loop + OpenCV) and Tracker.image_pyramid's
cv2.resize(INTER_NEAREST) on a seeded event text file.  Authoring container only:
    python tests/golden/make_event_golden.py
Stored: the event table (so the text file can be rebuilt), packet count / durations / mid-times, the signed
and unsigned frames of every packet and their 3-level pyramids.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GSEVT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))   # rosbag stub
sys.path.insert(0, REF)

import cv2  # noqa: E402
from utils.event_camera.event import EventFrame, load_events_from_txt  # noqa: E402

W, H, NPK = 96, 64, 3000
rng = np.random.default_rng(77)
n = 2 * NPK + 700                         # two full packets and a tail that must be dropped
ts = np.sort(rng.integers(10_000, 400_000, n))
x = rng.integers(0, W, n)
y = rng.integers(0, H, n)
p = rng.integers(0, 2, n)
x[:200], y[:200] = 0, 0                   # pile-ups on the corners / borders
x[200:300], y[200:300] = W - 1, H - 1
table = np.stack([ts, x, y, p], axis=1).astype(np.int64)
K = np.array([49.1, 0, 45.7, 0, 43.7, 31.4, 0, 0, 1.0]).reshape(3, 3)
D = np.array([-0.031982, 0.041966, -0.000507, -0.001031, 0.0])
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "events.txt")
    np.savetxt(path, table, fmt="%d", delimiter=" ")
    arrays = load_events_from_txt(path, NPK)
out = dict(table=table, W=W, H=H, NPK=NPK, K=K, D=D, n_packets=len(arrays),
           durations=np.array([a.duration() for a in arrays]), times=np.array([a.time() for a in arrays]))
for i, a in enumerate(arrays):
    ef = EventFrame(W, H, K, D, 9, a, device="cpu")
    s, u = ef.sign_delta_Ie.numpy(), ef.unsign_delta_Ie.numpy()
    out[f"sign_{i}"], out[f"unsign_{i}"] = s, u
    for l in range(3):
        sc = 0.5 ** l
        out[f"sign_{i}_L{l}"] = cv2.resize(s[0], (int(W * sc), int(H * sc)), interpolation=cv2.INTER_NEAREST)
np.savez_compressed(os.path.join(HERE, "events_96x64.npz"), **out)
print("packets", len(arrays), "durations", out["durations"], "cv2", cv2.__version__)
