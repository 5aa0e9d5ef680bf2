#!/usr/bin/env python
"""Debug: in-kernel mean vs torch for several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
for B in (1, 2, 3, 4, 5, 6, 8, 16, 32):
    for N in (1024, 512):
        fe = ri_b200.FrontEnd(B, N, 4, k=8, r=16, voxel_shape="cube")
        g = torch.Generator(device="cuda"); g.manual_seed(B * 7 + N)
        pts = torch.randn((B, 6, N), device="cuda", generator=g) * 5 - 1.3
        pts[:, :3, ::5] *= 300.0
        fe.load(pts, torch.randn((B, 4, N), device="cuda", generator=g))
        fe.forward(); torch.cuda.synchronize()
        same = torch.equal(fe._mean_buf, pts[:, :3, :].mean(2)) if fe._own_mean else None
        print("B=%2d N=%4d accepted=%s data-equal=%s" % (B, N, fe._own_mean, same))
