#!/usr/bin/env python
"""Timing of the hash-grid k-NN on ICL-NUIM-shaped scans (BASELINE configs[3]) next to the brute-force kernel."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
from ri_b200 import synth

B, N, k = 8, 50000, 20
x = torch.from_numpy(np.stack([synth.make_scan(N, seed=s)[:3] for s in range(B)])).cuda()
L = ri_b200._lib.lib
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((B, k, N), device="cuda"); i = torch.empty((B, k, N), dtype=torch.int32, device="cuda")
nws = L.ri_knn_grid_workspace_bytes(B, N, N); ws = torch.empty(nws, dtype=torch.uint8, device="cuda")


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = timeit(lambda: L.ri_knn_grid_f32(x.data_ptr(), x.data_ptr(), B, N, N, k, d.data_ptr(), i.data_ptr(), ws.data_ptr(), nws, st), 10)
b = timeit(lambda: L.ri_knn_f32(x.data_ptr(), x.data_ptr(), B, 3, N, N, k, d.data_ptr(), i.data_ptr(), st), 2)
print("k-NN %d scans x %d pts k=%d: grid %.3f ms (%.1f Mpts/s), brute force %.1f ms (%.2f Mpts/s), speed-up %.0fx"
      % (B, N, k, g, B * N / g / 1e3, b, B * N / b / 1e3, b / g))
