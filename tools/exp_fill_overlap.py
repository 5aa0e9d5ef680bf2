#!/usr/bin/env python
"""Grid writer: launches of consecutive batches on two streams (ramp and tail of one launch under the other), against one launch
after the other on one stream — what a launch costs in steady state.   python tools/exp_fill_overlap.py [sph|cube]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
shape = "spherical" if (len(sys.argv) < 2 or sys.argv[1] == "sph") else "cube"
B, N, C, k, r = 32, 1024, 67 if shape == "spherical" else 71, 20, 32
L = ri_b200._lib.lib
fes = []
for q in range(4):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, device="cuda")
    fe.load(synth.make_clouds(B, N, seed=1000 + q), synth.make_features(B, C, N, seed=1000 + q)); fe.forward(); fes.append(fe)
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def fill(fe, st):
    rc = L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st.cuda_stream)
    assert rc == 0


def run(nstreams, n=400):
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for s in streams:
        s.wait_event(e0)
    for i in range(n):
        fill(fes[i % 4], streams[i % nstreams])
    for s in streams:
        cur.wait_stream(s)
    e1.record(cur); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


run(1, 20); run(2, 20)
bytes_ = B * (C + 1) * r ** 3 * 4
for ns in (1, 2):
    us = run(ns)
    print("%s: %d stream(s): %.1f us per launch = %.2f TB/s" % (shape, ns, us, bytes_ / us / 1e6))
