#!/bin/bash
OUT=gpurun_out/r2o; mkdir -p $OUT
export RI_REQUIRE_REF=1
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin_reference_python.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for W in sph cube; do LANES=2 timeout 120 python tools/tune_lanes.py $W 2>&1 | tail -1 | tee -a $OUT/lanes.txt; done
SHAPE=sph LANES=1 timeout 120 python tools/timeline_step.py 2>&1 | tee $OUT/timeline_sph_serial.txt
