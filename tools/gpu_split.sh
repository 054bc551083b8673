#!/usr/bin/env bash
# Multi-GPU visit for the screen-tile split: gpurun --gpus N -- 'bash tools/gpu_split.sh <tag> <N>'
set -u
TAG="${1:-split}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/${TAG}_topo.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_tilesplit.py -q > "$OUT/${TAG}_pytest_tilesplit.log" 2>&1; tail -4 "$OUT/${TAG}_pytest_tilesplit.log"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29601 tools/tilesplit_check.py > "$OUT/${TAG}_check_n${N}.log" 2>&1; grep -v Warning "$OUT/${TAG}_check_n${N}.log" | tail -8
for cfg in "1000000 640 480" "5000000 1280 720"; do
  set -- $cfg
  for n in 1 $N; do
    if [ "$n" = 1 ]; then L="python"; else L="$TR --nproc-per-node $n --master-port 2961$n"; fi
    timeout 900 $L bench.py --gpus $n --mode tilesplit --gaussians $1 --width $2 --height $3 --steps 100 --warmup 5 \
      > "$OUT/${TAG}_bench_tilesplit_${1}_n${n}.json" 2> "$OUT/${TAG}_bench_tilesplit_${1}_n${n}.err"
    python - "$OUT/${TAG}_bench_tilesplit_${1}_n${n}.json" <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(r["config"]["workload"], "->", r["value"], "it/s", r["ms_per_step"], "ms", [(x["rows"], x["instances"]) for x in r["ranks"]])
except Exception as e:
    print("no result", sys.argv[1], e)
PY
  done
done
