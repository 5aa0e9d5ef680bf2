#!/bin/bash
# zero-stream grid writer against the ring form: equality, the writer alone, and the step in flight
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 600 python tools/exp_fill_form.py > $OUT/fill_form.json 2> $OUT/fill_form.err; echo "exp rc=$?"
cat $OUT/fill_form.json | tail -40
for W in sph cube; do
  for F in 1 2; do
    for LN in 2 3; do
      RI_FILL_FORM=$F LANES=$LN timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt
    done
  done
  RI_FILL_FORM=2 RI_FILL_WARPS=4 LANES=2 timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt
  RI_FILL_FORM=2 RI_FILL_CTAS=2 LANES=2 timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt
  RI_FILL_FORM=2 RI_FILL_WARPS=4 RI_FILL_CTAS=2 LANES=2 timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt
done
