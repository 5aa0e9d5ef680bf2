"""One line per configuration: the front-end step one at a time and three in flight, plus the grid writer alone.
   python tools/tune_lanes.py [sph|cube]   (RI_FILL_* / RI_* environment knobs select the variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
shape = "spherical" if (len(sys.argv) < 2 or sys.argv[1] == "sph") else "cube"
B, N, C, k, r = 32, 1024, 67 if shape == "spherical" else 71, 20, 32
L = ri_b200._lib.lib
NL = int(os.environ.get('LANES', 3))
fes = []
for q in range(max(3, NL)):
    fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=shape, device="cuda")
    fe.load(synth.make_clouds(B, N, seed=1000 + q), synth.make_features(B, C, N, seed=1000 + q)); fe.forward(); fes.append(fe)
torch.cuda.synchronize()
def timeit(fn, n=300):
    fn(12); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
serial = timeit(lambda n: [fes[i % len(fes)].forward() for i in range(n)])
lanes = ri_b200.FrontEndLanes(fes[:max(NL, 2)] if NL < 3 else fes, lanes=NL)
def lane_steps(n):
    lanes.begin()
    for i in range(n): lanes.forward(i)
    lanes.end()
inflight = timeit(lane_steps)
st = torch.cuda.current_stream().cuda_stream
def fill(n):
    for i in range(n):
        fe = fes[i % 3]
        L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
f = timeit(fill)
knobs = " ".join("%s=%s" % (k_, v) for k_, v in sorted(os.environ.items()) if k_.startswith("RI_") or k_ == "LANES")
print("%-9s serial %6.1f us  in flight %6.1f us  fill alone %5.1f us   %s" % (shape, serial, inflight, f, knobs))
