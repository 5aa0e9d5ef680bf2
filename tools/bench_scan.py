#!/usr/bin/env python
"""BASELINE configs[3] as one pipeline: ICL-NUIM-shaped raw scans (~400k points) -> barycentre grid subsampling to ~50k
(row f2) -> hash-grid k-NN k=20 (row a1') -> fused neighbour gather + PPF (a3) -> spherical voxelize r=64 + spherical devox
(a4/a5/a9).  CUDA events per stage, best of `REPS`; one JSON line.

    python tools/bench_scan.py [--scans 8] [--raw 400000] [--cell 0.045]"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
from ri_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=8); ap.add_argument("--raw", type=int, default=400000)
ap.add_argument("--cell", type=float, default=0.045); ap.add_argument("--k", type=int, default=20)
ap.add_argument("--res", type=int, default=64); ap.add_argument("--channels", type=int, default=16)
a = ap.parse_args()
REPS = 5
raw = [torch.from_numpy(np.ascontiguousarray(synth.make_scan(a.raw, seed=s).T)).cuda() for s in range(a.scans)]    # [N,6]


def timed(fn):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


# 1. subsample every scan (points + normals as the averaged "features", re-normalised afterwards)
def subsample():
    return [ri_b200.grid_sub_sampling(r[:, :3].contiguous(), features=r[:, 3:].contiguous(), grid_size=a.cell) for r in raw]
t_sub, subs = timed(subsample)


def subsample_many():                      # the same scans through the batched call: one host synchronisation for all of them
    return ri_b200.grid_sub_sampling_many([(r[:, :3].contiguous(), r[:, 3:].contiguous()) for r in raw], grid_size=a.cell)
t_sub_many, subs_many = timed(subsample_many)
assert all(torch.equal(p, q[0]) and torch.equal(f, q[1]) for (p, f), q in zip(subs, subs_many))
t_sub_one, t_sub = t_sub, t_sub_many
n = min(p.shape[0] for p, _ in subs)
xyz = torch.stack([p[:n].t().contiguous() for p, _ in subs]).contiguous()                      # [B,3,n]
nrm = torch.stack([torch.nn.functional.normalize(f[:n], dim=1).t().contiguous() for _, f in subs]).contiguous()
feat = torch.randn(a.scans, a.channels, n, device="cuda")
# 2. k-NN (hash grid, size-routed by the op), 3. PPF, 4. spherical voxelize + devox
t_knn, (dist, idx) = timed(lambda: ri_b200.functional.knn_indices(xyz, a.k))
t_ppf, ppf = timed(lambda: ri_b200.functional.knn_ppf(xyz, nrm, idx))
vox = ri_b200.modules.Spherical_Voxelization(a.res)
t_vox, (grid, ind, nc) = timed(lambda: vox(feat, xyz))
t_dev, dv = timed(lambda: ri_b200.functional.spherical_trilinear_devoxelize(grid, nc, ind, a.res))
total = t_sub + t_knn + t_ppf + t_vox + t_dev
print(json.dumps({"workload": "ICL-NUIM-shaped scans (BASELINE configs[3])", "scans": a.scans, "raw_points_per_scan": a.raw,
                  "cell_m": a.cell, "points_per_scan_after_subsampling": int(n), "k": a.k, "spherical_res": a.res,
                  "channels": a.channels,
                  "ms": {"grid_subsample": t_sub, "grid_subsample_one_call_per_scan": t_sub_one, "knn_hash_grid": t_knn, "ppf_gather": t_ppf, "sph_voxelize": t_vox,
                         "sph_devox": t_dev, "total": total},
                  "raw_points_per_s": a.scans * a.raw / (total * 1e-3),
                  "subsampled_points_per_s_knn_ppf_vox": a.scans * n / ((total - t_sub) * 1e-3),
                  "undefined_points": int((ind < 0).sum())}))
