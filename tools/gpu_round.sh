#!/usr/bin/env bash
# One GPU-box visit: parity tests, both bench arms, the ncu launch list of the bench command and one
# `ncu --set full` capture of every kernel of one iteration.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh v6 [tests] [bench] [ref] [launches] [full]'
set -u
TAG="${1:-run}"; shift || true
WHAT="${*:-tests bench ref launches full}"
OUT=gpurun_out
mkdir -p "$OUT"
has() { case " $WHAT " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/${TAG}_smi.txt" 2>&1
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_pytest_gpu.log" 2>&1
  tail -5 "$OUT/${TAG}_pytest_gpu.log"
fi
if has smoke; then
  timeout 300 python __graft_entry__.py smoke > "$OUT/${TAG}_smoke.log" 2>&1; tail -2 "$OUT/${TAG}_smoke.log"
fi
if has bench; then
  timeout 600 python bench.py > "$OUT/${TAG}_bench_ours.json" 2> "$OUT/${TAG}_bench_ours.err"
  cat "$OUT/${TAG}_bench_ours.json"; tail -3 "$OUT/${TAG}_bench_ours.err"
fi
if has ref; then
  timeout 600 python bench.py --impl reference > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
  cat "$OUT/${TAG}_bench_reference.json"; tail -3 "$OUT/${TAG}_bench_reference.err"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file "$OUT/${TAG}_launches.csv" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --no-parity-check \
      > "$OUT/${TAG}_launches.log" 2>&1
  tail -2 "$OUT/${TAG}_launches.log"
fi
if has full; then
  # kernels of the steady-state iterations only: skip the map packing / probing / warm-up launches
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:'preprocess_map|bucket_scatter|bucket_sort|blend_fwd|blend_bwd|geom_compact|geom_bwd|loss_stats|engine_update' \
      -s "${NCU_SKIP:-120}" -c "${NCU_COUNT:-40}" -o "$OUT/${TAG}_full" -f \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-parity-check > "$OUT/${TAG}_full.log" 2>&1
  ncu -i "$OUT/${TAG}_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_full_raw.csv" 2>/dev/null
  tail -2 "$OUT/${TAG}_full.log"
  ls -la "$OUT/${TAG}_full.ncu-rep"
fi
