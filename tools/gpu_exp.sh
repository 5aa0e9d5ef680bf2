#!/bin/bash
OUT=gpurun_out/r2r; mkdir -p $OUT
for W in sph cube; do
for V in "LANES=2" "LANES=2 RI_FE_KNN_FIRST=1" "LANES=3 RI_FE_KNN_FIRST=1" "LANES=2 RI_FE_KNN_FIRST=1 RI_FE_SIDE_PRIO=0"; do
  env $V timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt
done
done
