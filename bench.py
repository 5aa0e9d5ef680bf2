#!/usr/bin/env python
"""bench.py — feature-extractor front-end throughput (points/sec on 1024-point clouds) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cu_dg|sph_dg] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one pass of the front end (k-NN -> fused gather+PPF; coordinate prologue ->
voxelize -> trilinear devoxelize -> DGCNN voxel-neighbour edge features) over one batch of 32 synthetic
ModelNet40-shaped clouds of 1024 points PER GPU (weak scaling: clouds are sharded by rank, no data-path collective).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant HBM kernel (vox_fill: the dense [C, r^3] grid + count grid written exactly once), algorithmic
                bytes per launch / CUDA-event duration measured live, against MEASURED_PEAKS.json's HBM GB/s; the whole
                voxelize op and the whole step are quoted beside it
  cpu_baseline  the C oracle port (oracle/ri_oracle.c, OpenMP over clouds) timed on this box's host cores on a
                bounded sample of the same workload (rank 0, N=1 only)
  e2e           the same metric through the host-facing streaming call FrontEndPipeline.submit()/result(): every step
                copies its inputs pinned host -> device and its per-point outputs device -> pinned host; the copies of
                neighbouring steps overlap the compute (wall clock over K steps incl. the final drain); the
                one-call-at-a-time FrontEnd.run_staged() figure is quoted beside it
`--impl reference` times the UNMODIFIED reference kernels (oracle/_ref, the reference's own CUDA backend recompiled
for sm_100a — the reference has no CPU implementation: every op CHECK_CUDAs) through the reference's op sequence on
the same workload with the same host<->device copies; if that library is absent it times the oracle port on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "feature-extractor points/sec (1024-pt clouds)"
UNIT = "points/s"
WORKLOADS = {
    # BASELINE.json configs[1]: cu_dg variant (cube voxelization + DGCNN), 32 x 1024 on 1 B200; C = 3 + 4 + 64
    # input channels of the registration model's first PVConv
    "cu_dg": dict(voxel_shape="cube", B=32, N=1024, C=71, k=20, r=32,
                  name="cu_dg front end: 32 x 1024 pts with normals, k=20 KNN+PPF, cube voxelize r=32 C=71, "
                       "trilinear devox, DGCNN edge gather (BASELINE configs[1])"),
    # BASELINE.json configs[0]: sph_dg classification front end; C = 3 + 64
    "sph_dg": dict(voxel_shape="spherical", B=32, N=1024, C=67, k=20, r=32,
                   name="sph_dg front end: 32 x 1024 pts with normals, k=20 KNN+PPF, spherical voxelize r=32 C=67, "
                        "spherical trilinear devox, DGCNN edge gather (BASELINE configs[0])"),
}
# BASELINE.json configs[4]: throughput sweep over the spherical resolution (4096 clouds x 1024 pts = 128 steps of 32 clouds)
for _r in (16, 64):
    WORKLOADS["sph_r%d" % _r] = dict(voxel_shape="spherical", B=32, N=1024, C=67, k=20, r=_r,
                                     name="sph_dg front end at spherical res %d: 32 x 1024 pts per step, k=20 KNN+PPF, C=67 "
                                          "(BASELINE configs[4] sweep)" % _r)
RING = 3          # independent input/output buffer sets cycled between timed steps (footprint > L2)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the benchmark runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                try:
                    self.samples.append((time.time(), float(f[0]), float(f[1]), f[2:6]))
                except ValueError:
                    pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        note = "sampled inside the timed regions"
        if not inside:
            inside, note = self.samples, "timed regions shorter than the 100 ms sampling period: whole-run samples"
        mhz = sorted(s[1] for s in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in inside for n, v in zip(names, s[3]) if v.lower().startswith("active")})
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside), "note": note}


def traffic_from_profile(workload):
    """dram bytes per launch of the dominant op from the committed ncu capture, if one exists (else null)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def step_traffic_from_profile(workload):
    """dram bytes of ALL kernels of one step from the committed ncu capture (a lower bound of the HBM traffic), or null."""
    return traffic_from_profile(workload + "_step_total")


# ===================================================================================================== ours
def run_ours(args, rank, world, local):
    import numpy as np
    import torch
    import ri_b200
    from ri_b200 import shard, synth

    wl = WORKLOADS[args.workload]
    B, N, C, k, r = wl["B"], wl["N"], wl["C"], wl["k"], wl["r"]
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    sampler = ClockSampler(local) if rank == 0 else None

    # per-rank shard of the (world * B)-cloud job, RING independent batches
    engines, batches = [], []
    for q in range(RING):
        pts = synth.make_clouds(B, N, seed=1000 + 17 * rank + q)
        feats = synth.make_features(B, C, N, seed=1000 + 17 * rank + q)
        # throughput configuration (batches in flight): the per-cloud mean stays torch's kernel — with the mean fused into the
        # prefix kernel a single step is 7 us shorter but three steps in flight are 4 us per step slower (measured)
        fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=wl["voxel_shape"], normalize=False, device=dev, fuse_mean=False)
        fe.h_points.copy_(torch.from_numpy(pts)); fe.h_features.copy_(torch.from_numpy(feats))
        fe.load(fe.h_points, fe.h_features)
        engines.append(fe); batches.append((pts, feats))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    def timed(fn, steps):
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize(); barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev), (t0, time.time())

    windows = []
    # ---- device-resident throughput (`value`)
    #      RING steps in flight (FrontEndLanes: engine q replays on its own launch stream, so the latency-bound prefix
    #      and the k-NN of one batch run under the grid write / devoxelize of another); all K steps are launched after
    #      the start event and have finished before the end event.  The one-step-at-a-time figure is quoted beside it.
    for i in range(max(args.warmup, RING)):
        engines[i % RING].forward()
    # one step at a time: the latency configuration (mean fused into the prefix kernel)
    serial = []
    for q in range(RING):
        fe = ri_b200.FrontEnd(B, N, C, k=k, r=r, voxel_shape=wl["voxel_shape"], normalize=False, device=dev)
        fe.load(engines[q].h_points, engines[q].h_features); fe.forward(); serial.append(fe)
    torch.cuda.synchronize()
    ms_single, w = timed(lambda i: serial[i % RING].forward(), args.steps); windows.append(w)
    serial_fused_mean = bool(serial[0]._own_mean)
    del serial
    lanes = ri_b200.FrontEndLanes(engines, lanes=RING)

    def timed_lanes(steps):
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        lanes.begin()
        for i in range(steps):
            lanes.forward(i)
        lanes.end()
        e1.record()
        torch.cuda.synchronize(); barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev), (t0, time.time())
    timed_lanes(max(args.warmup, RING))
    ms, w = timed_lanes(args.steps); windows.append(w)
    pts_per_step = world * B * N
    value = pts_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the host-facing call (`e2e`): every step copies ITS inputs from pinned host memory to the
    #      device and ITS per-point outputs back to pinned host memory; the streaming API overlaps the copies of
    #      neighbouring steps with the compute (FrontEndPipeline), the one-call-at-a-time form is quoted beside it
    for i in range(max(3, min(args.warmup, 5))):
        engines[i % RING].run_staged()
    e2e_steps = max(1, min(args.steps, 200))
    ms_sync, w = timed(lambda i: engines[i % RING].run_staged(), e2e_steps); windows.append(w)
    pipe = ri_b200.FrontEndPipeline(B, N, C, depth=RING, k=k, r=r, voxel_shape=wl["voxel_shape"], normalize=False, device=dev)
    for q in range(RING):
        pipe.slot(q).h_points.copy_(torch.from_numpy(batches[q][0])); pipe.slot(q).h_features.copy_(torch.from_numpy(batches[q][1]))

    def pipe_step(i):
        pipe.submit(pipe.acquire())
    for i in range(2 * RING):
        pipe_step(i)
    pipe.drain()
    barrier(); torch.cuda.synchronize()
    t0 = time.time(); p0 = time.perf_counter()
    for i in range(e2e_steps):
        pipe_step(i)
    pipe.drain()
    torch.cuda.synchronize()
    ms_e2e = shard.max_over_ranks((time.perf_counter() - p0) * 1e3, dev); windows.append((t0, time.time()))
    barrier()
    e2e_value = pts_per_step * e2e_steps / (ms_e2e * 1e-3)
    del pipe

    # ---- dominant HBM kernel in isolation: vox_fill (the dense [C, r^3] grid + count grid written once), and the whole
    #      voxelize op (prefix + fill), CUDA events on the launching stream
    L = ri_b200._lib.lib
    st = torch.cuda.current_stream().cuda_stream
    shape_id = 2 if wl["voxel_shape"] == "spherical" else 0

    def fill_only(i):
        fe = engines[i % RING]
        rc = L.ri_voxelize_fill_f32(B, C, N, r, 0, B, fe.grid.data_ptr(), fe.cnt.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
        assert rc == 0

    def vox_op(i):
        fe = engines[i % RING]
        mean = fe._mean_buf if fe._own_mean else fe.points[:, :3, :].mean(2)
        rc = L.ri_vox_front_f32(fe.points.data_ptr(), 6, mean.data_ptr(), fe.features.data_ptr(), B, C, N, r, shape_id, 0.0,
                                fe.NORM_MODE | (0x100 if fe._own_mean else 0), fe.norm_coords.data_ptr(), fe._vox_coords.data_ptr(), fe.ind.data_ptr(),
                                fe.edge.data_ptr(), fe._ws.data_ptr(), fe._ws_bytes, st)
        assert rc == 0
        fill_only(i)
    for i in range(RING):
        vox_op(i)
    vox_steps = max(3, min(args.steps, 300))
    ms_fill, w = timed(fill_only, vox_steps); windows.append(w)
    ms_devox, w = timed(lambda i: engines[i % RING]._devox(0, B, st), vox_steps); windows.append(w)
    ms_vox, w = timed(vox_op, vox_steps); windows.append(w)
    alg = engines[0].algorithmic_bytes()
    peak, peak_src = measured_peaks()
    fill_bytes = B * (4 * (r ** 3) + 4 * C * (r ** 3))
    fill_gbs = fill_bytes / (ms_fill / vox_steps * 1e-3) / 1e9
    vox_gbs = alg["voxelize"] / (ms_vox / vox_steps * 1e-3) / 1e9
    step_gbs = alg["total"] / (ms / args.steps * 1e-3) / 1e9
    step_traffic = step_traffic_from_profile(args.workload)

    # ---- registration matcher (BASELINE configs[2]: 256 pairs x 1024 x 1024 x 512 over 8 GPUs = 32 pairs per GPU), an
    #      auxiliary figure: tcgen05 3xTF32 contraction + fused argmins, CUDA events, descriptors resident in HBM
    MP, MC, Mn = 32, 512, 1024
    g = torch.Generator(device=dev); g.manual_seed(4242 + rank)
    d1 = torch.randn((MP, MC, Mn), device=dev, generator=g); d2 = torch.randn((MP, MC, Mn), device=dev, generator=g)
    mm = ri_b200.matcher.MutualMatcher(MP, MC, Mn, Mn, device=dev)
    for _ in range(3):
        mm(d1, d2)
    m_steps = 20
    ms_match, w = timed(lambda i: mm(d1, d2), m_steps); windows.append(w)
    ms_match /= m_steps
    # ... and what follows it on the GPU (row f3): RANSAC (1000 hypotheses per pair) + Horn refit + RRE/RTE/RMSE
    src_p, tgt_p, Rg, tg = synth.make_pairs(MP, Mn, seed=77 + rank)
    p1 = torch.from_numpy(np.ascontiguousarray(src_p[:, :3].transpose(0, 2, 1))).to(dev)
    p2 = torch.from_numpy(np.ascontiguousarray(tgt_p[:, :3].transpose(0, 2, 1))).to(dev)
    ident = torch.arange(Mn, dtype=torch.int32, device=dev)[None].repeat(MP, 1).contiguous()
    cnt_p = torch.full((MP,), Mn, dtype=torch.int32, device=dev)
    gt_T = torch.eye(4, device=dev)[None].repeat(MP, 1, 1); gt_T[:, :3, :3] = torch.from_numpy(Rg).to(dev); gt_T[:, :3, 3] = torch.from_numpy(tg).to(dev)

    def pose_step(i):
        T, _ = ri_b200.registration.estimate_poses(p1, p2, ident, ident, cnt_p, func='ransac', seed=i)
        return ri_b200.registration.registration_metrics(gt_T, T, p1)
    for _ in range(3):
        pose_step(0)
    ms_pose, w = timed(pose_step, m_steps); windows.append(w)
    ms_pose /= m_steps
    pose_rre = float(pose_step(0)[:, 0].max())
    del mm, d1, d2

    if rank != 0:
        return
    sampler.stop()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "clouds_per_gpu": B, "points_per_cloud": N, "k": k, "resolution": r,
                   "channels": C, "voxel_shape": wl["voxel_shape"], "parallelism": "clouds sharded by rank (dp%d)" % world,
                   "l2": "inputs larger than L2: %d independent batches cycled, %.0f MB written per step" %
                         (RING, (alg["voxelize"] + alg["devox"] + alg["edge"] + alg["knn_ppf"]) / 1e6),
                   "cuda_graph": True,
                   "overlap": "k-NN/PPF branch on a side stream next to the grid writer and the devoxelizer; %d independent "
                              "batches in flight on %d launch streams" % (RING, RING),
                   "steps_in_flight": RING,
                   "one_step_at_a_time": {"ms_per_step": ms_single / args.steps,
                                          "value": pts_per_step * args.steps / (ms_single * 1e-3),
                                          "mean_fused_into_prefix_kernel": serial_fused_mean},
                   "grid_chunks": engines[0].grid_chunks},
        "roofline": {"bound": "hbm", "kernel": "vox_fill (dense [C,r^3] grid + count grid, written once)",
                     "achieved": fill_gbs, "peak": peak, "unit": "GB/s", "frac": fill_gbs / peak,
                     "traffic": traffic_from_profile(args.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": fill_bytes, "ms_per_launch": ms_fill / vox_steps,
                     "voxelize_op": {"kernels": ("vox_front (mean, prologue, cell sort, cell means, edge features) + vox_fill" if engines[0]._own_mean
                                                 else "torch mean + vox_front (prologue, cell sort, cell means, edge features) + vox_fill"),
                                     "algorithmic_bytes": alg["voxelize"], "ms": ms_vox / vox_steps,
                                     "achieved": vox_gbs, "frac": vox_gbs / peak},
                     "devoxelize_op": {"kernel": "devox_stream (planes streamed by TMA through a shared-memory ring)"
                                                 if not engines[0].join_before_devox else "devox (per-point gathers)",
                                       "algorithmic_bytes": alg["devox"], "ms": ms_devox / vox_steps,
                                       "achieved": alg["devox"] / (ms_devox / vox_steps * 1e-3) / 1e9,
                                       "frac": alg["devox"] / (ms_devox / vox_steps * 1e-3) / 1e9 / peak,
                                       "grid_bytes_read": B * 4 * C * r ** 3,
                                       "note": "algorithmic bytes count 32 B per point and channel (8 corners); at r=32, N=1024 "
                                               "the corners touch about every 32-byte sector, so the kernel reads the whole grid: "
                                               "grid_bytes_read / ms is its real HBM rate"},
                     "whole_step": {"algorithmic_bytes": alg["total"], "achieved": step_gbs, "frac": step_gbs / peak,
                                    "dram_traffic": step_traffic,
                                    "dram_traffic_gbs": (step_traffic / (ms / args.steps * 1e-3) / 1e9) if step_traffic else None,
                                    "dram_traffic_frac": (step_traffic / (ms / args.steps * 1e-3) / 1e9 / peak) if step_traffic else None,
                                    "note": "algorithmic bytes count the devoxelizer's 8 corners per point (32 B per point and "
                                            "channel); at sector granularity it reads the whole grid, which dram_traffic (sum of "
                                            "the kernels' dram bytes in the committed ncu capture, a lower bound) includes"}},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": engines[0].h2d_bytes,
                "d2h_bytes_per_step": engines[0].d2h_bytes, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "api": "FrontEndPipeline.submit/result (3 slots: H2D, step and D2H of neighbouring steps overlap)",
                "one_call_at_a_time": {"value": pts_per_step * e2e_steps / (ms_sync * 1e-3), "ms_per_step": ms_sync / e2e_steps,
                                       "api": "FrontEnd.run_staged()"}},
        "gpu_launches": engines[0].kernels_per_step * args.steps,
        "clocks": sampler.summary(windows),
        "matcher": {"workload": "mutual-NN matching, %d pairs x %d x %d x %d per GPU (BASELINE configs[2] at 8 GPUs)" % (MP, Mn, Mn, MC),
                    "pairs_per_s": world * MP / (ms_match * 1e-3), "ms_per_call": ms_match,
                    "useful_tflops": world * 2.0 * MP * Mn * Mn * MC / (ms_match * 1e-3) / 1e12,
                    "issued_tf32_tflops": world * 3 * 2.0 * MP * Mn * Mn * MC / (ms_match * 1e-3) / 1e12,
                    "note": "3xTF32 split precision: three tensor-core products per useful one; includes the re-tiling "
                            "pre-pass and the fp32 distance re-evaluation",
                    "pose": {"workload": "RANSAC 1000 hypotheses + Horn refit + RRE/RTE/RMSE, %d pairs x %d matches per GPU" % (MP, Mn),
                             "ms_per_call": ms_pose, "pairs_per_s": world * MP / (ms_pose * 1e-3), "max_rre_deg": pose_rre}},
    }
    if world == 1:
        line["cpu_baseline"] = cpu_port_baseline(wl, batches[0])
    print(json.dumps(line))


def cpu_port_step(wl, pts, feats, o):
    """The same front-end step with the C oracle port (OpenMP over clouds)."""
    import numpy as np
    B, N, C, k, r = pts.shape[0], wl["N"], wl["C"], wl["k"], wl["r"]
    xyz, nrm = np.ascontiguousarray(pts[:, :3]), np.ascontiguousarray(pts[:, 3:])
    _, idx = o.knn_one(xyz, xyz, k)
    gi = np.broadcast_to(idx.reshape(B, 1, k * N), (B, 3, k * N))
    o.ppf(np.broadcast_to(xyz[:, :, None, :], (B, 3, k, N)).reshape(B, 3, k * N), np.take_along_axis(xyz, gi, 2),
          np.broadcast_to(nrm[:, :, None, :], (B, 3, k, N)).reshape(B, 3, k * N), np.take_along_axis(nrm, gi, 2))
    if wl["voxel_shape"] == "spherical":
        avg, ind, nc = o.spherical_voxelization_module(feats, xyz, r)
        o.spherical_trilinear_devoxelize(nc, avg, ind, r)
    else:
        avg, ind, nc = o.voxelization_module(feats, xyz, r, normalize=False)
        o.trilinear_devoxelize(nc, avg, r)
    o.voxel_edge_gather(avg, feats, ind)


def cpu_port_baseline(wl, batch, reps=3):
    from oracle import cpu_oracle as o
    o.build()
    pts, feats = batch
    nb = min(pts.shape[0], 32)
    cpu_port_step(wl, pts[:2], feats[:2], o)              # warm-up (library load, page faults)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_port_step(wl, pts[:nb], feats[:nb], o)
        best = min(best, time.perf_counter() - t0)
    return {"value": nb * wl["N"] / best, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "%d clouds x %d pts, best of %d passes of the C oracle port (OpenMP over clouds, %d threads)" %
                      (nb, wl["N"], reps, os.cpu_count())}


# ================================================================================================ reference
def run_reference(args, rank, world, local):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    B, N, C, k, r = wl["B"], wl["N"], wl["C"], wl["k"], wl["r"]
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": wl["name"], "clouds_per_gpu": B, "points_per_cloud": N, "k": k,
                                            "resolution": r, "channels": C, "voxel_shape": wl["voxel_shape"]}}
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-registration-based-on-rotation-invariant-feature_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ri_synth", os.path.join(
        ROOT, "point-cloud-registration-based-on-rotation-invariant-feature_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    pts = synth.make_clouds(B, N, seed=1000); feats = synth.make_features(B, C, N, seed=1000)

    ref = None
    try:
        import torch
        from oracle.build_ref import load_ref
        if torch.cuda.is_available():
            ref = load_ref()
    except Exception:
        ref = None

    if ref is None:                                       # no reference library on this box: the oracle port
        cb = cpu_port_baseline(wl, (pts, feats), reps=max(1, min(args.steps, 3)))
        base.update({"value": cb["value"], "ms_per_step": B * N / cb["value"] * 1e3, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return

    import torch
    torch.cuda.set_device(0)
    dev = "cuda:0"
    hp = torch.from_numpy(pts).pin_memory(); hf = torch.from_numpy(feats).pin_memory()
    out_h = {}

    def step(host):
        """Reference op sequence for the same front end, through the reference backend's own functions
        (bindings.cpp:13-56) and the torch glue the reference uses around them."""
        p = hp.to(dev, non_blocking=True) if host else d_p
        f = hf.to(dev, non_blocking=True) if host else d_f
        xyz = p[:, :3].contiguous(); nrm = p[:, 3:].contiguous()
        _, _, idx, _ = ref.knn_forward_cuda(xyz, xyz, k)                                   # bilateral API
        gi = idx.reshape(B, 1, k * N).expand(-1, 3, -1).long()
        c_xyz = xyz[:, :, None, :].expand(-1, -1, k, -1).reshape(B, 3, k * N).contiguous()
        c_n = nrm[:, :, None, :].expand(-1, -1, k, -1).reshape(B, 3, k * N).contiguous()
        ppf = ref.spherical_ppf_forward(torch.gather(xyz, 2, gi), c_xyz, torch.gather(nrm, 2, gi), c_n)
        nc = xyz - xyz.mean(2, keepdim=True)
        if wl["voxel_shape"] == "spherical":
            nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values + 1e-20)
            avg, ind, cnt = ref.spherical_avg_voxelize_forward(f, nc.contiguous(), r)
            dv, _, _ = ref.spherical_trilinear_devoxelize_forward(r, True, nc.contiguous(), avg, ind)
        else:
            nc = torch.clamp((nc + 1) / 2.0 * r, 0, r - 1)
            avg, ind, cnt = ref.avg_voxelize_forward(f, torch.round(nc).to(torch.int32).contiguous(), r)
            dv, _, _ = ref.trilinear_devoxelize_forward(r, True, nc.contiguous(), avg)
        mask = ind == -1                                                                   # pvconv.py:68-90
        it = ind.clone(); it[mask] = 0
        centre = avg.gather(2, it.unsqueeze(1).expand(-1, C, -1).long())
        rel = f - centre
        rel[mask.unsqueeze(1).expand(-1, C, -1)] = 0
        edge = torch.cat((rel, f), 1)
        if host:
            for name, t in (("ppf", ppf), ("devox", dv), ("edge", edge)):
                if name not in out_h:
                    out_h[name] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                out_h[name].copy_(t, non_blocking=True)
            torch.cuda.synchronize()

    d_p, d_f = hp.to(dev), hf.to(dev)
    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    for _ in range(3):
        step(True)
    n2 = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(n2):
        step(True)
    ms2 = (time.perf_counter() - t0) * 1e3
    value = B * N * args.steps / (ms * 1e-3)
    base.update({
        "value": value, "ms_per_step": ms / args.steps,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": "the reference's own CUDA kernels (oracle/_ref, unmodified sources recompiled for "
                                   "sm_100a) on the same B200 — the reference has no CPU implementation of this path"},
        "e2e": {"value": B * N * n2 / (ms2 * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": hp.numel() * 4 + hf.numel() * 4,
                "d2h_bytes_per_step": sum(t.numel() * 4 for t in out_h.values()), "steps": n2,
                "ms_per_step": ms2 / n2}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cu_dg")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, local)
        return
    from ri_b200 import shard
    rank, world, local = shard.init_from_env()
    try:
        run_ours(args, rank, world, local)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
