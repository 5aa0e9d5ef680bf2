#!/usr/bin/env python
"""Timeline of the front-end step with several batches in flight (FrontEndLanes): GPU-timer stamps around every kernel of
every engine; prints the kernels of the last steps on one time axis (debug tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth, _lib
L, check = _lib.lib, _lib.check
B, N, C, k, r = 32, 1024, 71, 20, 32
NL = int(os.environ.get("LANES", 3))


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
        self._stamp("knn  <"); super()._knn(); self._stamp("knn  >")

    def _ppf(self):
        self._stamp("ppf  <"); super()._ppf(); self._stamp("ppf  >")

    def _devox(self, b0, b1, st):
        self._stamp("devox<"); super()._devox(b0, b1, st); self._stamp("devox>")

    def _branch_b(self, join=None, fork=None):
        st = torch.cuda.current_stream().cuda_stream
        self._stamp("mean <")
        mean = self.points[:, :3, :].mean(2)
        self._stamp("mean >")
        self._stamp("front<")
        check(L.ri_vox_front_f32(self.points.data_ptr(), 6, mean.data_ptr(), self.features.data_ptr(), B, C, N, r, 0, 0.0,
                                 self.NORM_MODE, self.norm_coords.data_ptr(), self._vox_coords.data_ptr(), self.ind.data_ptr(),
                                 self.edge.data_ptr(), self._ws.data_ptr(), self._ws_bytes, st), "front")
        self._stamp("front>")
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._knn(); self._ppf()
        self._stamp("fill <")
        check(L.ri_voxelize_fill_f32(B, C, N, r, 0, B, self.grid.data_ptr(), self.cnt.data_ptr(), self._ws.data_ptr(),
                                     self._ws_bytes, st), "fill")
        self._stamp("fill >")
        self._devox(0, B, st)


fes = []
for q in range(2 * NL):                      # the second half is not traced: it keeps the machine busy after the traced steps
    cls = Traced if q < NL else ri_b200.FrontEnd
    fe = cls(B, N, C, k=k, r=r, voxel_shape="cube")
    fe.load(synth.make_clouds(B, N, seed=q), synth.make_features(B, C, N, seed=q)); fes.append(fe)
ln = ri_b200.FrontEndLanes(fes, lanes=NL)
torch.cuda.synchronize()
ln.begin()
for i in range(8 * 2 * NL):
    ln.forward(i)
ln.end()
torch.cuda.synchronize()
rows = []
for e, fe in enumerate(fes[:NL]):
    t = fe.stamps.cpu().numpy()[:len(fe.names)]
    d = dict(zip(fe.names, t))
    for kname in ("mean ", "front", "knn  ", "ppf  ", "fill ", "devox"):
        rows.append((d[kname + "<"], d[kname + ">"], e, kname))
t0 = min(r_[0] for r_ in rows)
print("engine kernel   start     end    (us, last replay of each engine)")
for a, b, e, kname in sorted(rows):
    print("  %d    %s %8.1f %8.1f   %s%s" % (e, kname, (a - t0) / 1e3, (b - t0) / 1e3, " " * int((a - t0) / 4e3), "#" * max(1, int((b - a) / 4e3))))
