#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep — 4096 clouds x 1024 points through the sph_dg front end at spherical
resolution 16 / 32 / 64 (128 steps of 32 clouds per GPU-step), via bench.py.  Prints a markdown table.
    python tools/sweep.py [--gpus N]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gpus = int(sys.argv[sys.argv.index("--gpus") + 1]) if "--gpus" in sys.argv else 1
rows = []
for wl in ("sph_r16", "sph_dg", "sph_r64"):
    steps = 4096 // 32 // gpus
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(gpus), "--steps", str(steps), "--warmup", "5", "--workload", wl]
    if gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533"] + cmd[1:]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    j = json.loads(line)
    rows.append((j["config"]["resolution"], j["n_gpus"], j["value"], j["ms_per_step"], j["roofline"]["whole_step"]["frac"],
                 j["roofline"]["frac"], j["e2e"]["value"]))
print("| spherical res | GPUs | points/s (device-resident) | ms per 32-cloud step | whole-step HBM frac | vox_fill HBM frac | points/s end to end |")
print("|---|---|---|---|---|---|---|")
for r in rows:
    print("| %d | %d | %.3e | %.3f | %.2f | %.2f | %.3e |" % r)
