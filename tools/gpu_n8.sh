#!/usr/bin/env bash
# One visit with N GPUs (default 8): hypothesis-sharded bench (the driver's --gpus N line), the reference arm under torchrun,
# the tile-split check and the tile-split benches at N.   gpurun --gpus 8 -- 'bash tools/gpu_n8.sh <tag> 8'
set -u
TAG="${1:-n8}"; N="${2:-8}"
OUT=gpurun_out; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/${TAG}_topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 600 $TR --master-port 29701 bench.py --gpus $N --no-cpu-baseline > "$OUT/${TAG}_bench_hyp_n${N}.json" 2> "$OUT/${TAG}_bench_hyp_n${N}.err"
cut -c1-300 "$OUT/${TAG}_bench_hyp_n${N}.json"
timeout 600 $TR --master-port 29702 tools/tilesplit_check.py > "$OUT/${TAG}_check_n${N}.log" 2>&1; grep -v Warning "$OUT/${TAG}_check_n${N}.log" | tail -12
for cfg in "1000000 640 480" "5000000 1280 720"; do
  set -- $cfg
  timeout 900 $TR --master-port 29703 bench.py --gpus $N --mode tilesplit --gaussians $1 --width $2 --height $3 --steps 100 --warmup 5 \
      > "$OUT/${TAG}_bench_tilesplit_${1}_n${N}.json" 2> "$OUT/${TAG}_bench_tilesplit_${1}_n${N}.err"
  python - "$OUT/${TAG}_bench_tilesplit_${1}_n${N}.json" <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(r["config"]["workload"], "->", r["value"], "it/s", r["ms_per_step"], "ms", [(x["rows"], x["instances"]) for x in r["ranks"]])
    print(r["ranks"][0]["stages_ms"])
except Exception as e:
    print("no result", sys.argv[1], e)
PY
done
