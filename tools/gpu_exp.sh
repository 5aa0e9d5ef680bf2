#!/bin/bash
OUT=gpurun_out/r2tm; mkdir -p $OUT
for S in "2 512 1024 1024" "32 512 1024 1024" "32 512 1024 1024 reg" "3 100 300 700" "2 33 1000 132" "4 64 128 256" "5 128 1024 512 reg"; do
  timeout 90 python tools/exp_match_tma.py $S 2>&1 | tail -1 | tee -a $OUT/match_tma.txt
done
