#!/usr/bin/env python
"""Ball query + grouping at the shipped models' shape — B=32, N=M=1024, radius 0.3, u=128, grouping of
64-channel features ([B,64,1024,128] = 1.07 GB) — next to the reference's own kernels (oracle/_ref) on the same GPU."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
from ri_b200 import synth
B, N, U, C, R = 32, 1024, 128, 64, 0.3
pts = torch.from_numpy(np.ascontiguousarray(synth.make_clouds(B, N, seed=1)[:, :3])).cuda()
feat = torch.randn(B, C, N, device="cuda")

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

idx = torch.ops.ri.ball_query(pts, pts, R, U)
a = t(lambda: torch.ops.ri.ball_query(pts, pts, R, U))
b = t(lambda: torch.ops.ri.grouping(feat, idx))
out_bytes = B * C * N * U * 4
print("ours     : ball_query %8.1f us   grouping %8.1f us (%.0f GB/s of output written)" % (a, b, out_bytes / b / 1e3))
try:
    from oracle.build_ref import load_ref
    ref = load_ref()
    ra = t(lambda: ref.ball_query(pts, pts, R, U), 5)
    rb = t(lambda: ref.grouping_forward(feat, idx), 5)
    print("reference: ball_query %8.1f us   grouping %8.1f us   -> %.0fx / %.0fx" % (ra, rb, ra / a, rb / b))
    assert torch.equal(ref.ball_query(pts, pts, R, U), idx)
except Exception as ex:
    print("reference backend unavailable:", ex)

# the models' local PPF block: one kernel from the indices vs BallQuery grouping + the reference's torch ops (same GPU)
import ri_b200
nrm = torch.nn.functional.normalize(torch.randn(B, 3, N, device="cuda"), dim=1).contiguous()
grouper = ri_b200.modules.BallQuery(R, U, include_coordinates=True)


def torch_block():
    g = grouper(pts, pts, nrm)
    nc, nn_ = g[:, :3], g[:, 3:]
    ck = pts.unsqueeze(2).expand(-1, -1, U, -1); nk = nrm.unsqueeze(2).expand(-1, -1, U, -1)
    d = ck - nc
    dn = torch.norm(d, dim=1, p=2, keepdim=True)
    du = d / dn
    return torch.cat((torch.acos(nn_.mul(du).sum(1, keepdim=True).clamp(-1, 1)), torch.acos(nk.mul(du).sum(1, keepdim=True).clamp(-1, 1)),
                      torch.acos(nn_.mul(nk).sum(1, keepdim=True).clamp(-1, 1)), dn), 1)


c = t(lambda: ri_b200.functional.ball_local_ppf(pts, nrm, R, U))
d = t(torch_block, 5)
print("local PPF [B,4,U,N] incl. ball query: fused %8.1f us   BallQuery module + torch ops %8.1f us   -> %.1fx" % (c, d, d / c))

# the whole local-feature branch: indices -> PPF -> SharedMLP(4,[32,64]) -> max over the neighbours
fuser = ri_b200.modules.SharedMLP(4, [32, 64], dim=2).cuda().eval()
folded = ri_b200.functional.fold_fuser(fuser)
w1, b1, w2, b2 = folded


def torch_branch():
    with torch.no_grad():
        return fuser(torch_block()).max(dim=2).values


def torch_mlp_only(ppf):
    with torch.no_grad():
        return fuser(ppf).max(dim=2).values


ppf = ri_b200.functional.ball_local_ppf(pts, nrm, R, U)
e = t(lambda: torch.ops.ri.local_ppf_mlp_max(pts, nrm, pts, nrm, idx, w1, b1, w2, b2))
f = t(lambda: ri_b200.functional.local_ppf_features(pts, nrm, folded, R, U))
g = t(lambda: torch_mlp_only(ppf), 5)
h = t(torch_branch, 5)
flops = 2.0 * B * N * U * (4 * 32 + 32 * 64)
print("local-feature branch [B,64,N]: fused kernel alone %8.1f us (%.1f useful TFLOP/s), with ball query %8.1f us | torch layers on a "
      "ready PPF tensor %8.1f us, whole reference-style branch %8.1f us   -> %.1fx" % (e, flops / e / 1e6, f, g, h, h / f))
