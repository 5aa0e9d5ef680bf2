#!/usr/bin/env python
"""Timeline of one captured front-end step: GPU-timer stamps between the kernels of both branches (debug tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth, _lib
L, check = _lib.lib, _lib.check
B, N, C, k, r = 32, 1024, 71, 20, 32
shape = os.environ.get("SHAPE", "cube")

class Traced(ri_b200.FrontEnd):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.stamps = torch.zeros(64, dtype=torch.int64, device=self.device)
        self.names = []
    def _stamp(self, name):
        st = torch.cuda.current_stream().cuda_stream
        if name not in self.names:
            self.names.append(name)
        check(L.ri_debug_stamp(self.stamps.data_ptr() + 8 * self.names.index(name), st), "stamp")
    def _knn(self):
        self._stamp("A0 knn start")
        super()._knn()
        self._stamp("A1 knn done")
    def _ppf(self):
        self._stamp("A2 ppf start")
        super()._ppf()
        self._stamp("A3 ppf done")
    def _devox(self, b0, b1, st):
        self._stamp("B5 devox start")
        super()._devox(b0, b1, st)
        self._stamp("B6 devox done")
    def _branch_b(self, join=None, fork=None):
        self._stamp("B0 start")
        super()._branch_b(join=join, fork=fork)

fes = []
for q in range(3):
    fe = Traced(B, N, C, k=k, r=r, voxel_shape=shape)
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)); fe.forward(); fes.append(fe)
torch.cuda.synchronize()
for it in range(12):
    fes[it % 3].forward()
torch.cuda.synchronize()
for rep in range(2):
    fe = fes[rep]
    t = fe.stamps.cpu().numpy()[:len(fe.names)]
    t0 = t.min()
    print("--- replay", rep)
    for n, v in sorted(zip(fe.names, t), key=lambda x: x[1]):
        print("%-22s %8.1f us" % (n, (v - t0) / 1e3))
