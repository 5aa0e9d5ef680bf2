#!/usr/bin/env python
"""Experiment: stage-ordered pipelining.  Instead of one launch stream per batch (FrontEndLanes), one stream per STAGE
(prefix / grid write / devoxelize / k-NN+PPF): every stage runs its kernels in batch order, batches overlap only across
stages.  K steps over E engines are captured into one graph and replayed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth, _lib
L, check = _lib.lib, _lib.check
B, N, C, k, r = 32, 1024, int(os.environ.get("C", 71)), 20, 32
shape = os.environ.get("SHAPE", "cube")
E = int(os.environ.get("NE", 3)); K = int(os.environ.get("K", 24))
SPLIT_SIDE = os.environ.get("SPLIT_SIDE", "0") == "1"      # k-NN and PPF on separate stage streams
fes = []
for q in range(E):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, fuse_mean=os.environ.get("FUSE_MEAN", "0") == "1")
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)); fe.forward(); fes.append(fe)
torch.cuda.synchronize()
want = [{n: getattr(fe, n).clone() for n in ("ppf", "grid", "devox", "edge", "ind")} for fe in fes]
dev = fes[0].device
s_front, s_fill, s_devox, s_side, s_ppf = (torch.cuda.Stream(device=dev) for _ in range(5))
shape_id = 2 if shape == "spherical" else 0


def front(fe, st):
    own = fe._own_mean
    mean = fe._mean_buf if own else fe.points[:, :3, :].mean(2)
    check(L.ri_vox_front_f32(fe.points.data_ptr(), 6, mean.data_ptr(), fe.features.data_ptr(), B, C, N, r, shape_id, 0.0,
                             fe.NORM_MODE | (0x100 if own else 0), fe.norm_coords.data_ptr(), fe._vox_coords.data_ptr(),
                             fe.ind.data_ptr(), fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st), "front")


def fill(fe, st):
    check(L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st), "fill")


def build(K):
    """enqueue K steps on the stage streams (called under capture)"""
    cur = torch.cuda.current_stream()
    for s in (s_front, s_fill, s_devox, s_side, s_ppf):
        s.wait_stream(cur)
    done = [None] * E                      # per engine: events after which its buffers may be overwritten
    for i in range(K):
        fe = fes[i % E]
        with torch.cuda.stream(s_front):
            if done[i % E] is not None:
                for ev in done[i % E]: s_front.wait_event(ev)
            front(fe, s_front.cuda_stream)
            ev_front = torch.cuda.Event(); ev_front.record(s_front)
        with torch.cuda.stream(s_fill):
            s_fill.wait_event(ev_front)
            fill(fe, s_fill.cuda_stream)
            ev_fill = torch.cuda.Event(); ev_fill.record(s_fill)
        with torch.cuda.stream(s_devox):
            s_devox.wait_event(ev_fill)
            fe._devox(0, B, s_devox.cuda_stream)
            ev_devox = torch.cuda.Event(); ev_devox.record(s_devox)
        with torch.cuda.stream(s_side):
            s_side.wait_event(ev_front)
            fe._knn()
            if SPLIT_SIDE:
                ev_knn = torch.cuda.Event(); ev_knn.record(s_side)
            else:
                fe._ppf()
            ev_side = torch.cuda.Event(); ev_side.record(s_side)
        if SPLIT_SIDE:
            with torch.cuda.stream(s_ppf):
                s_ppf.wait_event(ev_knn)
                fe._ppf()
                ev_side = torch.cuda.Event(); ev_side.record(s_ppf)
        done[i % E] = (ev_devox, ev_side)
    for s in (s_front, s_fill, s_devox, s_side, s_ppf):
        cur.wait_stream(s)


warm = torch.cuda.Stream(device=dev)
with torch.cuda.stream(warm):
    build(E)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    build(K)
for _ in range(3): g.replay()
torch.cuda.synchronize()
ok = all(torch.equal(getattr(fe, n), w[n]) for fe, w in zip(fes, want) for n in w)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = max(1, 600 // K)
e0.record()
for _ in range(R): g.replay()
e1.record(); torch.cuda.synchronize()
print("stage-ordered: %d engines, %d steps per graph, split side %d: %7.1f us per step  (bits %s)" % (E, K, SPLIT_SIDE, e0.elapsed_time(e1) / (R * K) * 1e3, "equal" if ok else "DIFFER"))
