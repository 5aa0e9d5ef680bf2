#!/usr/bin/env python
"""BASELINE configs[2]: DeepGMR ModelNet40-Noisy-shaped registration — 256 source/target pairs x 1024 points, descriptors
512-d, pairs sharded over the GPUs of one box (both clouds of a pair on the same rank), per rank: tcgen05 mutual-NN matching
(row a11) -> RANSAC + Horn refit (row f3) -> RRE / RTE / RMSE; ONE NCCL all-gather of the results (poses + metrics, 76 bytes per
pair) at the end.  Descriptors are synthetic stand-ins for the extractor's output (the dense Conv3d / MLP layers are out of
scope): a fixed random projection of each point's coordinates IN THE SOURCE FRAME + noise, so that true partners match.

    python tools/bench_registration.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/bench_registration.py
"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
from ri_b200 import shard, synth

PAIRS, N, C = 256, 1024, 512
rank, world, local = shard.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
lo, hi = shard.shard_range(PAIRS, rank, world)
P = hi - lo
src, tgt, R, t = synth.make_pairs(PAIRS, N, seed=2024)                 # every rank draws the same job, keeps its slice
src, tgt, R, t = src[lo:hi], tgt[lo:hi], R[lo:hi], t[lo:hi]
g = torch.Generator(device=dev); g.manual_seed(7)
W = torch.randn((C, 3), device=dev, generator=g)
p1 = torch.from_numpy(np.ascontiguousarray(src[:, :3].transpose(0, 2, 1))).to(dev)       # [P,N,3]
p2 = torch.from_numpy(np.ascontiguousarray(tgt[:, :3].transpose(0, 2, 1))).to(dev)
Rt = torch.from_numpy(R).to(dev); tt = torch.from_numpy(t).to(dev)
back = torch.einsum("pij,pnj->pni", Rt.transpose(1, 2), p2 - tt[:, None, :])              # target points in the source frame
perm = torch.stack([torch.randperm(N, device=dev, generator=g) for _ in range(P)])
p2 = torch.gather(p2, 1, perm[:, :, None].expand(-1, -1, 3)).contiguous()                 # shuffle the target order
back = torch.gather(back, 1, perm[:, :, None].expand(-1, -1, 3))
f1 = torch.sin(torch.einsum("cj,pnj->pcn", W * 6.0, p1)).contiguous()                     # [P,C,N]
f2 = (torch.sin(torch.einsum("cj,pnj->pcn", W * 6.0, back)) + 0.02 * torch.randn((P, C, N), device=dev, generator=g)).contiguous()
gt = torch.eye(4, device=dev)[None].repeat(P, 1, 1); gt[:, :3, :3] = Rt; gt[:, :3, 3] = tt


def step():
    T, inl, m = ri_b200.registration.register_pairs(f1, f2, p1, p2, func="ransac")
    met = ri_b200.registration.registration_metrics(gt, T, p1)
    res = torch.cat([T.reshape(P, 16).double(), met, inl[:, None].double(), m["count"][:, None].double()], 1)   # [P,21]
    return shard.gather_clouds(res, PAIRS)                                                # the only collective


for _ in range(3):
    full = step()
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    full = step()
e1.record(); torch.cuda.synchronize()
ms = shard.max_over_ranks(e0.elapsed_time(e1) / K, dev)
if rank == 0:
    full = full.cpu().numpy()
    rre, rte, rmse, inl, cnt = full[:, 16], full[:, 17], full[:, 18], full[:, 19], full[:, 20]
    print(json.dumps({"workload": "DeepGMR-shaped registration (BASELINE configs[2]): %d pairs x %d pts, %d-d descriptors" % (PAIRS, N, C),
                      "n_gpus": world, "pairs_per_gpu": P, "ms_per_job": ms, "pairs_per_s": PAIRS / (ms * 1e-3),
                      "stages": "mutual-NN matcher (tcgen05 3xTF32) -> RANSAC 1000 hyp + Horn refit -> metrics -> all_gather(results)",
                      "gathered_bytes_per_pair": 21 * 8,
                      "mean_mutual_matches": float(cnt.mean()), "mean_inliers": float(inl.mean()),
                      "rre_deg_mean": float(rre.mean()), "rre_deg_max": float(rre.max()), "rte_mean": float(rte.mean()),
                      "rmse_mean": float(rmse.mean()), "recall_rmse_lt_0.2": float((rmse < 0.2).mean())}))
if world > 1:
    torch.distributed.destroy_process_group()
