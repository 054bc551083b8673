#!/usr/bin/env bash
# A/B of frame-0 tracking of the sequence test: this tree vs a copy of another build under _ab/ (scratch, git-ignored).
mkdir -p gpurun_out
O=$PWD/gpurun_out
for i in 1 2 3; do
  for side in old new; do
    if [ $side = new ]; then d=.; else d=_ab; fi
    (cd $d && timeout 200 python tests/test_gpu_sequence.py --frames 1 --out $O/seq_${side}_$i.json > $O/seq_${side}_$i.log 2>&1
     python - <<PY
import json
try:
    r=json.load(open("$O/seq_${side}_$i.json"))
    print("$side $i", r["per_frame_trans_m"], r["per_frame_rot_deg"], r["iterations_ours"], r["iterations_reference"])
except Exception as e:
    print("$side $i failed", e)
PY
    )
  done
done
