#!/usr/bin/env python
"""Experiment: the two forms of the cube devoxelizer (RI_DEVOX_STREAM=0 gathers / =1 TMA-streamed planes) alone, right after
the grid writer (grid partly in L2), next to the k-NN, and the whole step with / without the join in front of the devoxelizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
L = ri_b200._lib.lib
B, N, C, k, r = 32, 1024, int(os.environ.get("C", 71)), 20, int(os.environ.get("R", 32))
RING = 3
data = [(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)) for q in range(RING)]


def engines(**kw):
    fes = []
    for q in range(RING):
        fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape="cube", **kw)
        fe.load(*data[q]); fe.forward(); fes.append(fe)
    torch.cuda.synchronize()
    return fes


def timeit(fn, fes, n=60):
    for i in range(6): fn(fes[i % RING])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(fes[i % RING])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


st = torch.cuda.current_stream().cuda_stream
s2 = torch.cuda.Stream()
fes = engines(use_graph=False, overlap=False)


def devox(fe): fe._devox(0, B, st)
def fill(fe):
    assert L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def fill_devox(fe): fill(fe); devox(fe)
def knn(fe, stream=st):
    assert L.ri_knn_f32(fe.xyz.data_ptr(), fe.xyz.data_ptr(), B, 3, N, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), stream) == 0
def knn_devox(fe):
    cur = torch.cuda.current_stream(); s2.wait_stream(cur)
    with torch.cuda.stream(s2): knn(fe, s2.cuda_stream)
    devox(fe); cur.wait_stream(s2)
def knn_fill_devox(fe):
    cur = torch.cuda.current_stream(); s2.wait_stream(cur)
    with torch.cuda.stream(s2): knn(fe, s2.cuda_stream)
    fill(fe); devox(fe); cur.wait_stream(s2)


print("fill alone                : %7.1f us" % timeit(fill, fes))
print("knn alone                 : %7.1f us" % timeit(knn, fes))
print("devox alone               : %7.1f us" % timeit(devox, fes))
print("knn || (fill + devox)     : %7.1f us" % timeit(knn_fill_devox, fes))
del fes
for join in (True, False):
    g = engines(join_before_devox=join)
    print("step, join_before_devox=%-5s: %7.1f us" % (join, timeit(lambda fe: fe.forward(), g, n=200)))
    if not join:
        for nl in (1, 2, 3):
            ln = ri_b200.FrontEndLanes(g + engines(join_before_devox=join), lanes=nl)
            def run(n):
                ln.begin()
                for i in range(n): ln.forward(i)
                ln.end()
            run(12); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(300); e1.record(); torch.cuda.synchronize()
            print("  %d steps in flight: %7.1f us per step" % (nl, e0.elapsed_time(e1) / 300 * 1e3))
    del g
