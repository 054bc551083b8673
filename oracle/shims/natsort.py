"""TEST INFRASTRUCTURE — stand-in for `natsort` (reference utils/visualizer.py:5)."""
import re


def natsorted(seq):
    return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", str(s))])
