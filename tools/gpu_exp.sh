#!/bin/bash
OUT=gpurun_out/r2to; mkdir -p $OUT
for S in "32 512 1024 1024" "3 100 320 704" "2 33 1000 132" "5 128 1024 512 reg" "8 256 512 512"; do
  timeout 90 python tools/exp_match_tma.py $S 2>&1 | tail -1 | tee -a $OUT/match_tma.txt
done
