#!/usr/bin/env python
"""Tuning sweep for the front-end step: k-NN split factor, grid chunking, devox stream (CUDA events, rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import ri_b200
from ri_b200 import synth
L = ri_b200._lib.lib
B, N, C, k, r = 32, 1024, int(os.environ.get("C", 71)), 20, int(os.environ.get("R", 32))
shape = os.environ.get("SHAPE", "cube")
RING = 3
data = [(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)) for q in range(RING)]

def engines(**kw):
    fes = []
    for q in range(RING):
        fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, **kw)
        fe.load(*data[q]); fe.forward(); fes.append(fe)
        L.ri_split_xyz_normals_f32(fe.points.data_ptr(), B, N, fe.xyz.data_ptr(), fe.normals.data_ptr(), fe._packed.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return fes

def timeit(fn, fes, n=100):
    for i in range(6): fn(fes[i % RING])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(fes[i % RING])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

st = torch.cuda.current_stream().cuda_stream
s2 = torch.cuda.Stream()
fes = engines(use_graph=False, overlap=False, grid_chunks=1)
def knn(fe):
    assert L.ri_knn_f32(fe.xyz.data_ptr(), fe.xyz.data_ptr(), B, 3, N, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), st) == 0
print("knn          : %8.2f us" % timeit(knn, fes))
def ppf(fe):
    assert L.ri_ppf_gather_f32(fe.xyz.data_ptr(), fe.normals.data_ptr(), fe.knn_idx.data_ptr(), B, N, k, fe.ppf.data_ptr(), st) == 0
print("ppf gather   : %8.2f us" % timeit(ppf, fes))
def ppfp(fe):
    assert L.ri_ppf_gather_packed_f32(fe._packed.data_ptr(), fe.knn_idx.data_ptr(), B, N, k, fe.ppf.data_ptr(), st) == 0
print("ppf packed   : %8.2f us" % timeit(ppfp, fes))
def devox_ppf(fe):
    cur = torch.cuda.current_stream(); s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        assert L.ri_ppf_gather_packed_f32(fe._packed.data_ptr(), fe.knn_idx.data_ptr(), B, N, k, fe.ppf.data_ptr(), s2.cuda_stream) == 0
    fe._devox(0, B, st); cur.wait_stream(s2)
def knnppf(fe):
    assert L.ri_knn_ppf_f32(fe.points.data_ptr(), fe.points.data_ptr() + 3 * N * 4, 6 * N, B, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), fe.ppf.data_ptr(), st) == 0
print("knn+ppf fused: %8.2f us" % timeit(knnppf, fes))
sph = shape == "spherical"
def vox(fe):
    coords = fe.norm_coords if sph else fe._vox_coords
    fn = L.ri_sph_voxelize_edge_f32 if sph else L.ri_cube_voxelize_edge_f32
    assert fn(fe.features.data_ptr(), coords.data_ptr(), B, C, N, r, fe.grid.data_ptr(), fe.ind.data_ptr(), fe.cnt.data_ptr(), fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def vox_plain(fe):
    coords = fe.norm_coords if sph else fe._vox_coords
    fn = L.ri_sph_voxelize_f32 if sph else L.ri_cube_voxelize_f32
    assert fn(fe.features.data_ptr(), coords.data_ptr(), B, C, N, r, fe.grid.data_ptr(), fe.ind.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def prep(fe):
    coords = fe.norm_coords if sph else fe._vox_coords
    fn = L.ri_sph_voxelize_prepare_f32 if sph else L.ri_cube_voxelize_prepare_f32
    assert fn(coords.data_ptr(), B, C, N, r, fe.ind.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def means(fe):
    assert L.ri_voxelize_means_f32(fe.features.data_ptr(), B, C, N, r, 0, B, None, fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def means_edge(fe):
    assert L.ri_voxelize_means_f32(fe.features.data_ptr(), B, C, N, r, 0, B, fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def fill_only(fe):
    assert L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
def fill(fe):
    means(fe); fill_only(fe)
def fill_edge(fe):
    means_edge(fe); fill_only(fe)
# concurrency probe: the two branches' heavy kernels on two streams, eager
def both(fe):
    cur = torch.cuda.current_stream()
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        assert L.ri_knn_f32(fe.xyz.data_ptr(), fe.xyz.data_ptr(), B, 3, N, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), s2.cuda_stream) == 0
    fill(fe)
    cur.wait_stream(s2)
def both_devox(fe):
    cur = torch.cuda.current_stream()
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        assert L.ri_knn_f32(fe.xyz.data_ptr(), fe.xyz.data_ptr(), B, 3, N, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), s2.cuda_stream) == 0
    fe._devox(0, B, st)
    cur.wait_stream(s2)
print("voxelize (prepare+means+fill)      : %8.2f us" % timeit(vox_plain, fes))
print("voxelize+edge                      : %8.2f us" % timeit(vox, fes))
print("prepare only                       : %8.2f us" % timeit(prep, fes))
def front(fe):
    mean = fe.points[:, :3, :].mean(2)
    assert L.ri_vox_front_f32(fe.points.data_ptr(), 6, mean.data_ptr(), fe.features.data_ptr(), B, C, N, r, 2 if sph else 0, 0.0, 1, fe.norm_coords.data_ptr(), fe._vox_coords.data_ptr(), fe.ind.data_ptr(), fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
print("mean + fused front (+edge)         : %8.2f us" % timeit(front, fes))
print("means only                         : %8.2f us" % timeit(means, fes))
print("means+edge only                    : %8.2f us" % timeit(means_edge, fes))
print("fill only                          : %8.2f us" % timeit(fill_only, fes))
print("means+fill                         : %8.2f us" % timeit(fill, fes))
print("means(+edge)+fill                  : %8.2f us" % timeit(fill_edge, fes))
print("knn || means+fill (2 streams)      : %8.2f us" % timeit(both, fes))
print("knn || devox (2 streams)           : %8.2f us" % timeit(both_devox, fes))
print("devox || ppf packed (2 streams)    : %8.2f us" % timeit(devox_ppf, fes))
print("devox whole batch (grid cold)      : %8.2f us" % timeit(lambda fe: fe._devox(0, B, st), fes))
for after in (False, True):
    for join in (True, False):
        fes = engines(use_graph=True, overlap=True, knn_after_front=after, join_before_devox=join)
        print("step graph knn_after_front=%d join_before_devox=%d : %8.2f us" % (after, join, timeit(lambda fe: fe.forward(), fes, 200)))
        del fes
fes = engines(use_graph=True, overlap=False)
print("step graph serial : %8.2f us" % timeit(lambda fe: fe.forward(), fes, 200))
