#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 300 python tools/exp_devox_chans.py 2>&1 | tee $OUT/devox_chans.json
for CH in 8 12 16; do RI_DEVOX_CHANS=$CH LANES=2 timeout 120 python tools/tune_lanes.py sph 2>&1 | tail -1 | tee -a $OUT/lanes.txt; done
