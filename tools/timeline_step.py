#!/usr/bin/env python
"""Timeline of the production front-end step (the engine exactly as bench.py runs it) with LANES batches in flight: a
one-thread kernel that stores %globaltimer is enqueued before and after every library call of the step, on the stream the
call goes to, by wrapping the ctypes library object the engine calls through.  Prints the kernels of the last replay of each
traced engine on one time axis.

    SHAPE=sph|cube LANES=2 python tools/timeline_step.py      (RI_* knobs apply as usual)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth, _lib, frontend

SHAPE = "spherical" if os.environ.get("SHAPE", "sph") == "sph" else "cube"
NL = int(os.environ.get("LANES", 2))
B, N, C, k, r = 32, 1024, 67 if SHAPE == "spherical" else 71, 20, 32
TRACE = {"ri_vox_front_f32": "front", "ri_voxelize_fill_f32": "fill ", "ri_split_xyz_normals_f32": "split", "ri_knn_f32": "knn  ",
         "ri_ppf_gather_packed_f32": "ppf  ", "ri_sph_trilinear_devox_f32": "devox", "ri_trilinear_devox_f32": "devox"}
real = _lib.lib
cur = {"fe": None}


class StampLib:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if name not in TRACE:
            return fn
        label = TRACE[name]

        def wrapped(*a):
            fe = cur["fe"]
            if fe is None:
                return fn(*a)
            st = a[-1]
            for tag in ("<", ">"):
                if label + tag not in fe.names:
                    fe.names.append(label + tag)
            real.ri_debug_stamp(fe.stamps.data_ptr() + 8 * fe.names.index(label + "<"), st)
            rc = fn(*a)
            real.ri_debug_stamp(fe.stamps.data_ptr() + 8 * fe.names.index(label + ">"), st)
            return rc
        return wrapped


frontend._L = StampLib()
fes = []
for q in range(max(3, 2 * NL)):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=SHAPE, normalize=False)
    fe.stamps = torch.zeros(64, dtype=torch.int64, device=fe.device); fe.names = []
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q))
    cur["fe"] = fe if q < NL else None           # the others are not traced: they keep the machine busy after the traced steps
    fe.forward()                                 # verify + capture with the stamps inside the graph
    torch.cuda.synchronize()
    fes.append(fe)
cur["fe"] = None
ln = ri_b200.FrontEndLanes(fes, lanes=NL)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 8 * len(fes)
ln.begin(); e0.record()
for i in range(steps):
    ln.forward(i)
ln.end(); e1.record()
torch.cuda.synchronize()
print("%s, %d in flight: %.1f us per step (with the stamp kernels in the graphs)" % (SHAPE, NL, e0.elapsed_time(e1) / steps * 1e3))
rows = []
for e, fe in enumerate(fes[:NL]):
    t = fe.stamps.cpu().numpy()[:len(fe.names)]
    d = dict(zip(fe.names, t))
    for label in sorted(set(TRACE.values())):
        if label + "<" in d:
            rows.append((d[label + "<"], d[label + ">"], e, label))
t0 = min(r_[0] for r_ in rows)
print("engine kernel   start     end    (us, last replay of each traced engine)")
for a, b, e, label in sorted(rows):
    print("  %d    %s %8.1f %8.1f   %s%s" % (e, label, (a - t0) / 1e3, (b - t0) / 1e3, " " * int((a - t0) / 4e3), "#" * max(1, int((b - a) / 4e3))))
