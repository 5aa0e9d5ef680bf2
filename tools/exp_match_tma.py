#!/usr/bin/env python
"""Indices-only matcher: tensor-map (no image) path against the pre-pass path — same matches? time?
    python tools/exp_match_tma.py P C n1 n2"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
L = ri_b200._lib.lib
P, C, n1, n2 = [int(x) for x in sys.argv[1:5]]
g = torch.Generator(device="cuda"); g.manual_seed(7 + C)
d1 = torch.randn((P, C, n1), device="cuda", generator=g); d2 = torch.randn((P, C, n2), device="cuda", generator=g)
if len(sys.argv) > 5 and sys.argv[5] == "reg":                       # registration-shaped: d2 = permuted d1 + noise
    perm = torch.stack([torch.randperm(n1, device="cuda", generator=g) for _ in range(P)])
    d2 = torch.gather(d1, 2, perm[:, None, :].expand(-1, C, -1))[:, :, :n2].contiguous() + 0.05 * torch.randn((P, C, n2), device="cuda", generator=g)
mm = ri_b200.matcher.MutualMatcher(P, C, n1, n2, want_dist=False)


def t(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


assert L.ri_debug_set_knob(b"RI_MATCH_TMA", 0) == 0
mm(d1, d2); torch.cuda.synchronize()
ref = {k: getattr(mm, k).clone() for k in ("corr12", "corr21", "idx1", "idx2", "count")}
us_prep = t(lambda: mm(d1, d2))
out = {"shape": [P, C, n1, n2], "us_prepass_path": us_prep}
for mode, label in ((3, "3d_boxes"), (-1, "default")):
    assert L.ri_debug_set_knob(b"RI_MATCH_TMA", mode) == 0
    for k in ("corr12", "corr21", "idx1", "idx2", "count"):
        getattr(mm, k).fill_(-5)
    mm(d1, d2); torch.cuda.synchronize()
    got = {k: getattr(mm, k).clone() for k in ref}
    out["us_tensor_map_" + label] = t(lambda: mm(d1, d2))
    out["equal_" + label] = all(bool(torch.equal(ref[k], got[k])) for k in ref)
print(json.dumps(out))
