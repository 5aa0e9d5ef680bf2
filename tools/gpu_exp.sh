#!/bin/bash
OUT=gpurun_out/r2h; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --only > $OUT/n2_default.json 2>$OUT/n2_default.err
NCCL_MAX_CTAS=1 timeout 300 $TR --master-port 29542 bench.py --gpus 2 --only > $OUT/n2_maxctas1.json 2>$OUT/n2_maxctas1.err
NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 timeout 300 $TR --master-port 29543 bench.py --gpus 2 --only > $OUT/n2_nch1.json 2>$OUT/n2_nch1.err
for f in $OUT/n2_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; done
