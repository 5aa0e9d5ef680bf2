#!/bin/bash
OUT=gpurun_out/r2t; mkdir -p $OUT
export RI_REQUIRE_REF=1
timeout 600 python -m pytest tests/test_knn_warp_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "knn" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for W in sph cube; do LANES=2 timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt; done
timeout 100 python tools/bench_knn.py 2>&1 | tail -3
