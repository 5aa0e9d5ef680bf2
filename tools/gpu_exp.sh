#!/bin/bash
OUT=gpurun_out/r2tp; mkdir -p $OUT
export RI_REQUIRE_REF=1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
