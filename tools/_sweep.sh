mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tilesplit.py -q > gpurun_out/v7_tilesplit.log 2>&1; tail -15 gpurun_out/v7_tilesplit.log
i=0
for cfg in "--gaussians 8000 --scale-mult 2.5 --events 12000" "--gaussians 20000 --scale-mult 2.0 --events 12000" "--gaussians 20000 --scale-mult 3.0 --events 30000" "--gaussians 8000 --scale-mult 4.0 --events 30000 --lin-scale 1.0"; do
  i=$((i+1))
  timeout 400 python tests/test_gpu_sequence.py --frames 6 $cfg --out gpurun_out/seq_sweep_$i.json > gpurun_out/seq_sweep_$i.log 2>&1
  echo "cfg $i: $cfg"; python - <<PY
import json
r=json.load(open("gpurun_out/seq_sweep_$i.json"))
print(r["per_frame_trans_m"], r["per_frame_rot_deg"]); print("ours gt", r["unaligned_ours_vs_gt"]); print("ref gt", r["unaligned_reference_vs_gt"]); print(r["iterations_ours"], r["iterations_reference"])
PY
done
