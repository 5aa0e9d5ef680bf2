"""Where the end-to-end step loses time against the link bound (tools/exp_pcie.py): pipeline depth, copies without compute,
compute without copies.   python tools/exp_e2e.py [sph|cube]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth

shape = "spherical" if (len(sys.argv) < 2 or sys.argv[1] == "sph") else "cube"
B, N, C, k, r = 32, 1024, 67 if shape == "spherical" else 71, 20, 32
pts = torch.from_numpy(synth.make_clouds(B, N, seed=1)); feats = torch.from_numpy(synth.make_features(B, C, N, seed=1))


def run(depth, steps=400, mode="full"):
    pipe = ri_b200.FrontEndPipeline(B, N, C, depth=depth, k=k, r=r, voxel_shape=shape, device="cuda", edge_echo=False)
    for q in range(depth):
        pipe.slot(q).h_points.copy_(pts); pipe.slot(q).h_features.copy_(feats)
    if mode == "copies":
        for fe in pipe.slots:
            fe.forward = lambda: None
    if mode == "compute":
        for fe in pipe.slots:
            fe.h2d_off = True
    for i in range(2 * depth):
        pipe.submit(pipe.acquire())
    pipe.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        pipe.submit(pipe.acquire())
    pipe.drain(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    del pipe
    return ms


for depth in (2, 3, 4, 6):
    print("depth %d: full %.3f ms/step   copies only %.3f ms/step" % (depth, run(depth), run(depth, mode="copies")))
