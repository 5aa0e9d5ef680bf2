"""Per-op timings (CUDA events, rotating buffers) for the front-end kernels at the benchmark configuration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import ri_b200
from ri_b200 import synth
L = ri_b200._lib.lib
B, N, C, k, r = 32, 1024, int(os.environ.get("C", 71)), 20, int(os.environ.get("R", 32))
shape = os.environ.get("SHAPE", "cube")
RING = 3
fes = []
for q in range(RING):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, use_graph=False, overlap=False)
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q))
    fe.forward()
    fes.append(fe)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream

def timeit(name, fn, n=60):
    for i in range(6): fn(fes[i % RING])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(fes[i % RING])
    e1.record(); torch.cuda.synchronize()
    print("%-34s %8.2f us" % (name, e0.elapsed_time(e1) / n * 1e3))

sph = shape == "spherical"
def vox(fe, edge):
    coords = fe.norm_coords if sph else fe._vox_coords
    if edge:
        fn = L.ri_sph_voxelize_edge_f32 if sph else L.ri_cube_voxelize_edge_f32
        rc = fn(fe.features.data_ptr(), coords.data_ptr(), B, C, N, r, fe.grid.data_ptr(), fe.ind.data_ptr(), fe.cnt.data_ptr(), fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
    else:
        fn = L.ri_sph_voxelize_f32 if sph else L.ri_cube_voxelize_f32
        rc = fn(fe.features.data_ptr(), coords.data_ptr(), B, C, N, r, fe.grid.data_ptr(), fe.ind.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
    assert rc == 0
def devox(fe):
    if sph:
        rc = L.ri_sph_trilinear_devox_f32(fe.norm_coords.data_ptr(), fe.grid.data_ptr(), fe.ind.data_ptr(), B, C, N, r, fe.devox.data_ptr(), fe.devox_inds.data_ptr(), fe.devox_wgts.data_ptr(), st)
    else:
        rc = L.ri_trilinear_devox_f32(fe.norm_coords.data_ptr(), fe.grid.data_ptr(), B, C, N, r, fe.devox.data_ptr(), fe.devox_inds.data_ptr(), fe.devox_wgts.data_ptr(), st)
    assert rc == 0
def knn(fe):
    assert L.ri_knn_f32(fe.xyz.data_ptr(), fe.xyz.data_ptr(), B, 3, N, N, k, fe.knn_dist.data_ptr(), fe.knn_idx.data_ptr(), st) == 0
def ppf(fe):
    assert L.ri_ppf_gather_f32(fe.xyz.data_ptr(), fe.normals.data_ptr(), fe.knn_idx.data_ptr(), B, N, k, fe.ppf.data_ptr(), st) == 0
def edge(fe):
    assert L.ri_voxel_edge_gather_f32(fe.grid.data_ptr(), fe.features.data_ptr(), fe.ind.data_ptr(), B, C, N, r ** 3, fe.edge.data_ptr(), st) == 0
def fill_then_devox(fe):
    vox(fe, True); devox(fe)

print("config: B=%d N=%d C=%d k=%d r=%d %s" % (B, N, C, k, r, shape))
timeit("voxelize (prepare+fill)", lambda fe: vox(fe, False))
timeit("voxelize+edge (prepare+fill)", lambda fe: vox(fe, True))
timeit("devox (grid cold)", devox)
timeit("voxelize+edge then devox", fill_then_devox)
timeit("edge gather standalone", edge)
timeit("knn", knn)
timeit("ppf gather", ppf)
for fe in fes: fe.overlap = False
timeit("step eager, serial", lambda fe: fe._step())
for fe in fes: fe.overlap = True
timeit("step eager, overlap", lambda fe: fe._step())
for fe in fes: fe.use_graph = True
timeit("step graph, overlap", lambda fe: fe.forward())
