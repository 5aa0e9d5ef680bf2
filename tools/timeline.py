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
    def _branch_a(self):
        st = torch.cuda.current_stream().cuda_stream
        B, N, k = self.B, self.N, self.k
        self._stamp("A0 start")
        check(L.ri_split_xyz_normals_f32(self.points.data_ptr(), B, N, self.xyz.data_ptr(), self.normals.data_ptr(), st), "split")
        check(L.ri_knn_f32(self.xyz.data_ptr(), self.xyz.data_ptr(), B, 3, N, N, k, self.knn_dist.data_ptr(), self.knn_idx.data_ptr(), st), "knn")
        self._stamp("A2 knn done")
        check(L.ri_ppf_gather_f32(self.xyz.data_ptr(), self.normals.data_ptr(), self.knn_idx.data_ptr(), B, N, k, self.ppf.data_ptr(), st), "ppf")
        self._stamp("A3 ppf done")
    def _branch_b(self, join=None):
        st = torch.cuda.current_stream().cuda_stream
        B, N, C, r = self.B, self.N, self.C, self.r
        self._stamp("B0 start")
        mean = self.points[:, :3, :].mean(2)
        self._stamp("B1 mean done")
        check(L.ri_vox_front_f32(self.points.data_ptr(), 6, mean.data_ptr(), self.features.data_ptr(), B, C, N, r, 0, 0.0, 1,
                                 self.norm_coords.data_ptr(), self._vox_coords.data_ptr(), self.ind.data_ptr(), self.edge.data_ptr(),
                                 self._ws.data_ptr(), self._ws_bytes, st), "front")
        self._stamp("B3 front done")
        check(L.ri_voxelize_fill_f32(B, C, N, r, 0, B, self.grid.data_ptr(), self.cnt.data_ptr(), self._ws.data_ptr(), self._ws_bytes, st), "fill")
        self._stamp("B4 fill done")
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        self._stamp("B5 joined A")
        self._devox(0, B, st)
        self._stamp("B6 devox done")

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
