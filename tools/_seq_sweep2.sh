#!/usr/bin/env bash
# ours-only sweep of synthetic-sequence generators: is there one the tracker holds?
mkdir -p gpurun_out
i=0
for cfg in "--gaussians 40000 --width 320 --height 240 --events 12000" \
           "--gaussians 20000 --width 320 --height 240 --events 30000 --scale-mult 2.0" \
           "--gaussians 5000 --width 320 --height 240 --events 60000 --scale-mult 4.0" \
           "--gaussians 2000 --width 320 --height 240 --events 60000 --scale-mult 6.0 --ang-scale 2 --lin-scale 1" \
           "--gaussians 20000 --width 640 --height 480 --events 100000 --scale-mult 3.0 --ang-scale 2 --lin-scale 1" \
           "--gaussians 300000 --width 640 --height 480 --events 30000 --ang-scale 1 --lin-scale 1"; do
  i=$((i+1))
  timeout 300 python tools/long_sequence.py --frames 40 $cfg --out gpurun_out/sweep2_$i.json > /dev/null 2> gpurun_out/sweep2_$i.err
  echo "cfg $i: $cfg"
  python - <<PY
import json
try:
    r=json.load(open("gpurun_out/sweep2_$i.json"))
    print("  trans err every 5th frame:", r["trans_err_m_every_5th_frame"])
    print("  rot err:", r["rot_err_deg_every_5th_frame"], "iters/level mean", r["iterations_per_level_mean"], "max", r["iterations_per_level_max"])
except Exception as e:
    print("  failed", e)
PY
done
