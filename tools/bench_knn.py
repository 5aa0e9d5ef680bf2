"""k-NN alone on the bench workload (32 x 1024 self query, k = 20): the thread-per-query kernel vs the warp-per-query kernel.
   python tools/bench_knn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
from ri_b200 import synth
L = ri_b200._lib.lib

B, N, k = 32, 1024, 20
dev = torch.device("cuda")
sets = []
for q in range(3):
    x = torch.from_numpy(synth.make_clouds(B, N, seed=1000 + q)[:, :3].copy()).to(dev).contiguous()
    sets.append((x, torch.empty((B, k, N), device=dev), torch.empty((B, k, N), dtype=torch.int32, device=dev)))
st = torch.cuda.current_stream().cuda_stream


def warp(i):
    x, d, ix = sets[i % 3]
    assert L.ri_knn_f32(x.data_ptr(), x.data_ptr(), B, 3, N, N, k, d.data_ptr(), ix.data_ptr(), st) == 0


def brute(i):
    x, d, ix = sets[i % 3]
    assert L.ri_knn_thread_f32(x.data_ptr(), x.data_ptr(), B, 3, N, N, k, d.data_ptr(), ix.data_ptr(), st) == 0


def timeit(fn, steps=200):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


brute(0); torch.cuda.synchronize()
ref = [(s[1].clone(), s[2].clone()) for s in sets[:1]]
warp(0); torch.cuda.synchronize()
print("warp-per-query equal to brute force:", torch.equal(ref[0][1], sets[0][2]) and torch.equal(ref[0][0], sets[0][1]))
print("brute  %.1f us" % timeit(brute))
print("warp   %.1f us" % timeit(warp))
