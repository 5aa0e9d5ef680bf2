#!/usr/bin/env python
"""Experiment: steps in flight (FrontEndLanes) against step time, cu_dg front end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
B, N, C, k, r = 32, 1024, int(os.environ.get("C", 71)), 20, int(os.environ.get("R", 32))
shape = os.environ.get("SHAPE", "cube")
NE = 6
fes = []
for q in range(NE):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, join_before_devox={"0": False, "1": True}.get(os.environ.get("JOIN", ""), None),
                          knn_after_front=os.environ.get("KNN_AFTER", "1") == "1", fuse_mean=os.environ.get("FUSE_MEAN", "1") == "1", overlap=os.environ.get("OVERLAP", "1") == "1")
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)); fes.append(fe)
torch.cuda.synchronize()
for ne in (int(os.environ.get("NE", 3)),):
    for nl in (1, ne):
        ln = ri_b200.FrontEndLanes(fes[:ne], lanes=nl)
        def run(n):
            ln.begin()
            for i in range(n): ln.forward(i)
            ln.end()
        run(12); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(300); e1.record(); torch.cuda.synchronize()
        print("%d engines, %d launch streams: %7.1f us per step" % (ne, nl, e0.elapsed_time(e1) / 300 * 1e3))
