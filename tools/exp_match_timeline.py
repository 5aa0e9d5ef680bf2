#!/usr/bin/env python
"""Debug: %globaltimer stamps of CTA 0 of the persistent matcher GEMM (RI_MATCH_DBG=1 makes the kernel write them into the
'mutual' scratch array, which match_dist overwrites afterwards — so the GEMM is run alone here through a private copy)."""
import os, sys, ctypes
os.environ["RI_MATCH_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ri_b200
P, C, n = 32, int(os.environ.get("C", 512)), 1024
d1 = torch.randn((P, C, n), device="cuda"); d2 = torch.randn((P, C, n), device="cuda")
mm = ri_b200.matcher.MutualMatcher(P, C, n, n)
for _ in range(3): mm(d1, d2)
torch.cuda.synchronize()
# the stamps live at the start of the 'mutual' region; match_dist rewrote it with flags (0/1), so re-derive by reading
# before match_dist is impossible from here: instead read them from the workspace tail the GEMM wrote LAST (ints 0/1 << stamps)
ws = mm._ws
L = ri_b200._lib.lib
# layout: mutual region offset = total - P*n1p*4 ; recompute like match_ws_layout
n1p = n2p = (n + 255) // 256 * 256; Cp = (C + 15) // 16 * 16
o = 2 * P * n1p * Cp * 4 + 2 * P * n1p * 4; o = (o + 15) // 16 * 16; o += 2 * P * n1p * 8
base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
raw = ws[base + o: base + o + 64 * 8].view(torch.int64).cpu().numpy().reshape(8, 8)
t0 = raw[raw > 0].min()
names = ["mma: tile reached", "mma: buffer free", "mma: tile issued", "epi: ready to wait", "epi: accumulator full", "epi: TMEM read done", "epi: tile done"]
for it in range(8):
    print("tile %d: " % it + "  ".join("%s %.1f" % (names[e], (raw[it, e] - t0) / 1e3) for e in range(7) if raw[it, e] > 0))
