#!/usr/bin/env bash
# ncu --set full of the binning / projection kernels of two steady-state iterations; exports CSV pages on the box and drops
# the report when it is too large to travel back (gpurun_out is limited to 64 MiB).
TAG="${1:-run}"; OUT=gpurun_out; mkdir -p $OUT
REGEX="${2:-preprocess_map|bucket_scatter|bucket_sort|geom_compact}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s "${NCU_SKIP:-40}" -c "${NCU_COUNT:-14}" \
    -o "$OUT/${TAG}_full" -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_full.log" 2>&1
tail -2 "$OUT/${TAG}_full.log"
ncu -i "$OUT/${TAG}_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_full_raw.csv" 2>/dev/null
ncu -i "$OUT/${TAG}_full.ncu-rep" --page source --csv > "$OUT/${TAG}_full_source.csv" 2>/dev/null
ls -la $OUT/${TAG}_full*
sz=$(stat -c %s "$OUT/${TAG}_full.ncu-rep")
if [ "$sz" -gt 30000000 ]; then rm -f "$OUT/${TAG}_full.ncu-rep"; echo "report dropped ($sz bytes), CSV pages kept"; fi
