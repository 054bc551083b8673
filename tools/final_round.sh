#!/usr/bin/env bash
# Final verification of a build on ONE B200: the whole GPU suite, compute-sanitizer on the paths added in round 2
# (bulk id staging, split projection kernels / list scatter), smoke, both bench arms.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1800 -- 'bash tools/final_round.sh r2final'
set -u
TAG="${1:-final}"
OUT=gpurun_out
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/${TAG}_smi.txt" 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > "$OUT/${TAG}_pytest_gpu.log" 2>&1
grep -E "passed|failed" "$OUT/${TAG}_pytest_gpu.log" | tail -2
{
  echo "== racecheck: bulk id staging (GSEVT_BLEND_BULK=1), smoke()"
  GSEVT_BLEND_BULK=1 timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py smoke 2>&1 | tail -3
  echo "== memcheck: bulk id staging test"
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "optional_paths and BULK" 2>&1 | tail -4
  # (the sanitizer serialises kernels, so ranks that wait for each other cannot run under it: the missing-peer test drives
  # ONE rank of a 2-way split through the split projection kernels, the list scatter, the sort and the blend forward)
  echo "== memcheck: screen-tile split kernels (pre-test, list projection, list scatter), one rank of two"
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tilesplit.py -m gpu -x -q -k "missing_peer" 2>&1 | tail -4
} > "$OUT/${TAG}_sanitizer.log" 2>&1
grep -E "==|ERROR SUMMARY|passed|failed|hazards" "$OUT/${TAG}_sanitizer.log"
timeout 300 python __graft_entry__.py smoke > "$OUT/${TAG}_smoke.log" 2>&1; tail -1 "$OUT/${TAG}_smoke.log"
timeout 600 python bench.py > "$OUT/${TAG}_bench_ours.json" 2> "$OUT/${TAG}_bench_ours.err"
python - "$OUT/${TAG}_bench_ours.json" <<'PY'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("ours", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "parity", j["parity_check"]["ok"], "search", j["search"].get("hypotheses_per_s"),
      "traffic", j["roofline"]["traffic"])
PY
timeout 600 python bench.py --impl reference > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
python -c "import json,sys; j=json.load(open('$OUT/${TAG}_bench_reference.json')); print('reference', j['value'], j['ms_per_step'])"
