"""The grid writer alone on the bench workloads (RI_FILL_RING = slots per warp, RI_FILL_CTAS = CTAs per SM).
   python tools/bench_fill.py [sph|cube] [r]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
shape = "spherical" if (len(sys.argv) < 2 or sys.argv[1] == "sph") else "cube"
r = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B, N, C, k = 32, 1024, 67 if shape == "spherical" else 71, 20
L = ri_b200._lib.lib
fes = []
for q in range(3):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, device="cuda")
    fe.load(synth.make_clouds(B, N, seed=1000 + q), synth.make_features(B, C, N, seed=1000 + q)); fe.forward(); fes.append(fe)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
def fill(i):
    fe = fes[i % 3]
    assert L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st) == 0
ref = fes[0].grid.clone(); refc = fes[0].cnt.clone()
fes[0].grid.fill_(7.0); fes[0].cnt.fill_(7)
fill(0); torch.cuda.synchronize()
ok = torch.equal(ref, fes[0].grid) and torch.equal(refc, fes[0].cnt)
for i in range(10): fill(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(300): fill(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 300 * 1e3
byt = B * (C + 1) * r ** 3 * 4
print("%s r=%d: fill %.1f us  %.0f GB/s  rewrite identical: %s   (ring=%s ctas=%s)" % (shape, r, us, byt / us / 1e3, ok,
      os.environ.get("RI_FILL_RING", "-"), os.environ.get("RI_FILL_CTAS", "-")))
