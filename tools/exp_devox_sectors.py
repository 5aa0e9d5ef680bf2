"""How much of a grid plane does the trilinear devoxelizer NEED?  Counts, per cloud, the distinct 32-byte sectors (and 128-byte
lines, and x-slabs) that the 8 corners of its points touch in one channel plane, on the bench clouds.
   python tools/exp_devox_sectors.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth

B, N, k = 32, 1024, 20
for shape, C, r in (("cube", 71, 32), ("spherical", 67, 32), ("cube", 71, 16), ("cube", 71, 64)):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, device="cuda")
    fe.load(synth.make_clouds(B, N, seed=1000), synth.make_features(B, C, N, seed=1000)); fe.forward()
    torch.cuda.synchronize()
    idx = fe.devox_inds.long()                                  # [B,8,N] flat cell index of every corner
    w = fe.devox_wgts
    used = idx[w != 0] if shape == "cube" else idx              # a corner with weight 0 is still loaded by the reference
    s = r ** 3
    rows = []
    for b in range(B):
        cells = torch.unique(idx[b].reshape(-1))
        sect = torch.unique(cells // 8); line = torch.unique(cells // 32); slab = torch.unique(cells // (r * r))
        rows.append((cells.numel() / s, sect.numel() / (s / 8), line.numel() / (s / 32), slab.numel() / r))
    t = torch.tensor(rows).mean(0).tolist()
    occ = float((fe.cnt > 0).float().mean())
    print("%-9s r=%2d: occupied cells %.1f %% | corners touch %.1f %% of the cells, %.1f %% of the 32-byte sectors, %.1f %% of the "
          "128-byte lines, %.0f %% of the x-slabs of a plane  -> compulsory read %.1f MB of a %.1f MB grid (B=%d, C=%d)"
          % (shape, r, 100 * occ, 100 * t[0], 100 * t[1], 100 * t[2], 100 * t[3], t[1] * B * C * s * 4 / 1e6, B * C * s * 4 / 1e6, B, C))
    u = sorted(int(round(x[1] * s / 8)) for x in rows)
    print("          distinct sectors per cloud: min %d median %d max %d" % (u[0], u[len(u) // 2], u[-1]))
