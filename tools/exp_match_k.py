#!/usr/bin/env python
"""Experiment: matcher kernels at several channel counts (run under ncu --metrics gpu__time_duration.sum to split
per-tile fixed cost from per-stage cost)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ri_b200
P, n = 32, 1024
for C in (128, 256, 512, 1024):
    d1 = torch.randn((P, C, n), device="cuda"); d2 = torch.randn((P, C, n), device="cuda")
    mm = ri_b200.matcher.MutualMatcher(P, C, n, n)
    for _ in range(3):
        mm(d1, d2)
    torch.cuda.synchronize()
