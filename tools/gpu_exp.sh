#!/bin/bash
OUT=gpurun_out/r2m; mkdir -p $OUT
export RI_BENCH_MIN_MS=0
timeout 600 ncu --set full --clock-control none --import-source on -s 8 -c 2 -k regex:'vox_front' -o $OUT/front -f \
  python bench.py --only --workload sph_dg --steps 2 --warmup 3 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
