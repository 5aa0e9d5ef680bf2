#!/usr/bin/env python
"""Debug: the streaming devoxelizer alone, with and without its gather work (RI_DEVOX_DBG_SKIP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
B, N, C, r = 32, 1024, 71, 32
L = ri_b200._lib.lib
st = torch.cuda.current_stream().cuda_stream
bufs = []
for q in range(3):
    nc = (torch.rand(B, 3, N, device="cuda") * (r - 1)).contiguous()
    grid = torch.randn(B, C, r, r, r, device="cuda")
    bufs.append((nc, grid, torch.empty(B, C, N, device="cuda"), torch.empty(B, 8, N, dtype=torch.int32, device="cuda"), torch.empty(B, 8, N, device="cuda")))
def run(i):
    nc, g, o, di, dw = bufs[i % 3]
    assert L.ri_trilinear_devox_f32(nc.data_ptr(), g.data_ptr(), B, C, N, r, o.data_ptr(), di.data_ptr(), dw.data_ptr(), st) == 0
for mode in (None, "1"):
    if mode: os.environ["RI_DEVOX_DBG_SKIP"] = mode    # read once per process: run one mode per process
    for i in range(6): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60): run(i)
    e1.record(); torch.cuda.synchronize()
    print("skip gathers" if mode else "full        ", "%.1f us" % (e0.elapsed_time(e1) / 60 * 1e3))
