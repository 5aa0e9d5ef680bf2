#!/usr/bin/env python
"""Grid writer, ring form (RI_FILL_FORM=1) against the zero-stream form (RI_FILL_FORM=2): the two must write identical
grids (checked on several shapes, the zero-stream form repeated to catch ordering races between its bulk zero stores and
its patch stores), then both are timed alone (CUDA events) over a few launch shapes.

    python tools/exp_fill_form.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200
L = ri_b200._lib.lib


def knob(name, v):
    assert L.ri_debug_set_knob(name.encode(), int(v)) == 0, name


def engine(B, N, C, r, shape):
    pts = ri_b200.synth.make_clouds(B, N, seed=5)
    feats = ri_b200.synth.make_features(B, C, N, seed=5)
    fe = ri_b200.FrontEnd(B, N, C, k=8, r=r, voxel_shape=shape, device="cuda:0", use_graph=False)
    fe.load(torch.from_numpy(pts), torch.from_numpy(feats))
    return fe


def fill(fe):
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ri_voxelize_fill_f32(fe.B, fe.C, fe.N, fe.r, 0, fe.B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
    assert rc == 0, rc


def timed(fn, n=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = {"equal": {}, "us": {}}
ok = True
for (B, N, C, r, shape) in [(32, 1024, 67, 32, "spherical"), (32, 1024, 71, 32, "cube"), (8, 1024, 16, 16, "spherical"),
                            (4, 1024, 9, 64, "cube"), (3, 600, 5, 8, "cube"), (5, 1000, 3, 12, "spherical"), (2, 128, 1, 4, "cube")]:
    knob("RI_FILL_FORM", 1)
    fe = engine(B, N, C, r, shape)
    fe.forward(); torch.cuda.synchronize()
    g1, c1 = fe.grid.clone(), fe.cnt.clone()
    knob("RI_FILL_FORM", 2)
    same = True
    for rep in range(30):
        fe.grid.fill_(float("nan")); fe.cnt.fill_(-7)
        fill(fe); torch.cuda.synchronize()
        same = same and bool(torch.equal(fe.grid, g1)) and bool(torch.equal(fe.cnt, c1))
    # and inside the whole step, other kernels running next to it
    for rep in range(10):
        fe.grid.fill_(float("nan")); fe.cnt.fill_(-7)
        fe.forward(); torch.cuda.synchronize()
        same = same and bool(torch.equal(fe.grid, g1)) and bool(torch.equal(fe.cnt, c1))
    out["equal"]["%s B%d N%d C%d r%d" % (shape, B, N, C, r)] = same
    ok = ok and same
    if B == 32:
        key = "%s r%d C%d" % (shape, r, C)
        knob("RI_FILL_FORM", 1)
        out["us"][key + " ring"] = timed(lambda: fill(fe))
        knob("RI_FILL_FORM", 2)
        for warps in (4, 8):
            for ctas in (1, 2):
                knob("RI_FILL_WARPS", warps); knob("RI_FILL_CTAS", ctas)
                out["us"][key + " zero-stream %dw x %dcta" % (warps, ctas)] = timed(lambda: fill(fe))
        knob("RI_FILL_WARPS", -1); knob("RI_FILL_CTAS", -1)
    del fe
# r = 64 at full batch (2.3 GB grid)
knob("RI_FILL_FORM", 1)
fe = engine(32, 1024, 67, 64, "spherical")
fe.forward(); torch.cuda.synchronize()
g1 = fe.grid.clone()
out["us"]["spherical r64 C67 ring"] = timed(lambda: fill(fe), 30)
knob("RI_FILL_FORM", 2)
fill(fe); torch.cuda.synchronize()
out["equal"]["spherical B32 r64"] = bool(torch.equal(fe.grid, g1))
ok = ok and out["equal"]["spherical B32 r64"]
out["us"]["spherical r64 C67 zero-stream"] = timed(lambda: fill(fe), 30)
knob("RI_FILL_FORM", -1)
out["all_equal"] = ok
print(json.dumps(out, indent=1))
