#!/usr/bin/env bash
# Round-end evidence in one visit: all GPU tests, smoke, both bench arms, the ncu launch list of the bench command and one
# `ncu --set full` capture of two steady-state iterations exported as CSV pages (the report itself is dropped when it is
# too large for gpurun_out's 64 MiB).   gpurun --timeout 1800 -- 'bash tools/final_round.sh <tag>'
set -u
TAG="${1:-final}"; OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_round.sh "$TAG" tests smoke bench ref launches
REGEX='preprocess_map|bucket_scatter|bucket_sort|blend_fwd|blend_bwd|geom_compact|geom_bwd|loss_stats|engine_update'
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s "${NCU_SKIP:-100}" -c "${NCU_COUNT:-40}" \
    -o "$OUT/${TAG}_full" -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_full.log" 2>&1
tail -2 "$OUT/${TAG}_full.log"
ncu -i "$OUT/${TAG}_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_full_raw.csv" 2>/dev/null
ncu -i "$OUT/${TAG}_full.ncu-rep" --page details --csv > "$OUT/${TAG}_full_details.csv" 2>/dev/null
for k in blend_fwd_kernelILi1ELb0 blend_bwd_kernelILi1ELb0 bucket_sort_kernel bucket_scatter_kernel preprocess_map_kernelILi3; do
  python tools/ncu_lines.py "$OUT/${TAG}_full.ncu-rep" $k --top 30 > "$OUT/${TAG}_lines_$k.txt" 2>&1
done
ls -la $OUT/${TAG}_full*
sz=$(stat -c %s "$OUT/${TAG}_full.ncu-rep")
if [ "$sz" -gt 25000000 ]; then rm -f "$OUT/${TAG}_full.ncu-rep"; echo "report dropped ($sz bytes), CSV pages kept"; fi
du -sh $OUT
