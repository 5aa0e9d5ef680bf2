#!/usr/bin/env python
"""Bring-up / timing tool for the tcgen05 matcher (not a test): parity against an fp64 evaluation for a ladder of
shapes, then throughput at the DeepGMR shape.   python tools/bringup_matcher.py [--time]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ri_b200  # noqa: E402


def check(P, C, n1, n2, point_major, seed=0, structured=False):
    g = np.random.default_rng(seed)
    f1 = g.standard_normal((P, n1, C)).astype(np.float32)
    if structured and n1 == n2:
        f2 = np.stack([f1[p][g.permutation(n1)] for p in range(P)]) + 0.05 * g.standard_normal((P, n2, C)).astype(np.float32)
        f2 = f2.astype(np.float32)
    else:
        f2 = g.standard_normal((P, n2, C)).astype(np.float32)
    a = torch.from_numpy(f1 if point_major else np.ascontiguousarray(f1.transpose(0, 2, 1))).cuda()
    b = torch.from_numpy(f2 if point_major else np.ascontiguousarray(f2.transpose(0, 2, 1))).cuda()
    r = ri_b200.matcher.mutual_nn(a, b, point_major=point_major)
    torch.cuda.synchronize()
    bad12 = bad21 = 0
    worst = 0.0
    for p in range(P):
        x, y = f1[p].astype(np.float64), f2[p].astype(np.float64)
        d = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2 * x @ y.T
        scale = (x * x).sum(1).max() + (y * y).sum(1).max()
        c12 = r['corr12'][p].cpu().numpy(); c21 = r['corr21'][p].cpu().numpy()
        # a pick is acceptable when its fp64 distance is within 1e-5 * scale of the true minimum
        bad12 += int((d[np.arange(n1), c12] - d.min(1) > 1e-5 * scale).sum())
        bad21 += int((d[c21, np.arange(n2)] - d.min(0) > 1e-5 * scale).sum())
        worst = max(worst, float(np.abs(r['dist12'][p].cpu().numpy() - d[np.arange(n1), c12]).max() / scale))
        exact12 = int((c12 != d.argmin(1)).sum()); exact21 = int((c21 != d.argmin(0)).sum())
    print("P=%d C=%d n1=%d n2=%d pm=%d struct=%d : bad12=%d bad21=%d (last pair exact-mismatch %d/%d) dist_err=%.2e count=%s"
          % (P, C, n1, n2, point_major, structured, bad12, bad21, exact12, exact21, worst, r['count'][:4].tolist()))
    return bad12 + bad21


def main():
    tot = 0
    for args in [(1, 16, 128, 256, False), (1, 16, 128, 256, True), (1, 64, 128, 256, False), (1, 512, 128, 256, False),
                 (1, 512, 1024, 1024, False), (2, 512, 1024, 1024, True), (3, 100, 300, 700, False), (2, 33, 1000, 130, True)]:
        tot += check(*args)
    tot += check(2, 512, 1024, 1024, False, structured=True)
    print("TOTAL BAD", tot)
    if "--time" in sys.argv:
        P, C, n = 256, 512, 1024
        a = torch.randn(P, C, n, device="cuda"); b = torch.randn(P, C, n, device="cuda")
        m = ri_b200.matcher.MutualMatcher(P, C, n, n)
        for _ in range(3):
            m(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(a, b)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("matcher %d pairs x %d x %d x %d: %.3f ms/call, %.1f useful TFLOP/s, %.1f issued (x3) TFLOP/s, %.0f pairs/s"
              % (P, n, n, C, ms, m.flops / ms / 1e9, 3 * m.flops / ms / 1e9, P / ms * 1e3))
        t0 = time.perf_counter()
        x = a[0].T.cpu().numpy(); y = b[0].T.cpu().numpy()
        from oracle import cpu_oracle
        t0 = time.perf_counter()
        for _ in range(5):
            cpu_oracle.find_correspondence_one_pair(x, y)
        print("numpy reference: %.2f ms/pair" % ((time.perf_counter() - t0) / 5 * 1e3))


if __name__ == "__main__":
    main()
