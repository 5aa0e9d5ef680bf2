#!/bin/bash
# 8-GPU box: host-link table (tools/exp_pcie.py) at 1/2/4/8 ranks with and without NUMA binding, then both bench arms at N = 8
OUT=gpurun_out/r2s; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $OUT/lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 python tools/exp_pcie.py > $OUT/pcie_n1.json 2>$OUT/pcie_n1.err
for N in 2 4 8; do
  timeout 180 $TR --nproc-per-node $N --master-port $((29500+N)) tools/exp_pcie.py > $OUT/pcie_n$N.json 2>$OUT/pcie_n$N.err
done
RI_NO_BIND=1 timeout 180 $TR --nproc-per-node 8 --master-port 29520 tools/exp_pcie.py > $OUT/pcie_n8_nobind.json 2>$OUT/pcie_n8_nobind.err
timeout 400 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --only > $OUT/bench_n8.json 2>$OUT/bench_n8.err; echo "bench n8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --impl reference > $OUT/bench_ref_n8.json 2>$OUT/bench_ref_n8.err; echo "ref n8 rc=$?"
for f in $OUT/pcie_n*.json; do echo $f; tail -1 $f; done
tail -c 600 $OUT/bench_n8.json
