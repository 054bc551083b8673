#!/usr/bin/env python3
"""DRAM traffic per engine stage from an `ncu --set full` raw CSV -> profiles/ncu_traffic.json.

    ncu -i gpurun_out/<tag>_full.ncu-rep --page raw --csv > profiles/<name>_raw.csv
    python tools/ncu_traffic.py profiles/<name>_raw.csv [--out profiles/ncu_traffic.json]

bench.py reads the JSON to fill `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch of
the stage; a stage made of several library kernels, e.g. a CUB radix sort, is the sum of its kernels).  The
numbers are a one-off capture of the same workload (bench.py defaults), not measured during the timed run.
"""
import argparse
import collections
import csv
import json
import os

STAGE_OF = [  # (substring of the kernel name, stage)
    ("preprocess_map_kernel", "preprocess_map"), ("bucket_scatter_kernel", "bucket_scatter"), ("bucket_sort_kernel", "bucket_sort"),
    ("blend_fwd_kernel", "blend_fwd_gray"), ("loss_stats_kernel", "loss_stats"), ("blend_bwd_kernel", "blend_bwd_gray"),
    ("geom_compact_kernel", "geom_bwd_pose"), ("geom_bwd_kernel", "geom_bwd_pose"), ("engine_update_kernel", "engine_update"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    k, r, w, t = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
    sr, sw = scale[units[r]], scale[units[w]]
    per = collections.defaultdict(lambda: [0.0, 0.0, 0])   # bytes, us, iterations seen
    seen_in_iter = collections.defaultdict(int)
    after = None   # last non-CUB stage seen: CUB kernels after preprocess belong to the depth sort / scan, after emit to the tile sort
    # only complete iterations: from a projection kernel to the next update kernel
    body = rows[2:]
    starts = [i for i, row in enumerate(body) if "preprocess_map_kernel" in row[k]]
    keep = []
    for si in starts:
        end = next((j for j in range(si, len(body)) if "engine_update_kernel" in body[j][k]), None)
        if end is not None:
            keep.extend(body[si:end + 1])
    for row in keep:
        name = row[k]
        stage = next((s for sub, s in STAGE_OF if sub in name), None)
        if stage is None:
            continue
        b = float(row[r]) * sr + float(row[w]) * sw
        per[stage][0] += b
        per[stage][1] += float(row[t])
        if "preprocess_map_kernel" in name:
            seen_in_iter["iters"] += 1
    iters = max(1, seen_in_iter["iters"])
    out = {"source": os.path.basename(a.raw_csv), "iterations_captured": iters,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of each engine stage (ncu --set full, bench.py default workload)",
           "stages": {s: {"traffic_bytes": round(v[0] / iters), "ncu_us": round(v[1] / iters, 2)} for s, v in per.items()}}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
