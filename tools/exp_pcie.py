"""Host link of one rank: pinned H2D alone, D2H alone, both at once (the front end's end-to-end step moves 9.6 MB up and
28 MB down per 32 clouds), plus the NUMA placement shard.bind_host_to_gpu() finds.  One process per GPU under torchrun
shows what the ranks of a box do to each other:
    python tools/exp_pcie.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/exp_pcie.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ri_b200 import shard

rank, world, local = shard.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
bind = None if os.environ.get("RI_NO_BIND") else shard.bind_host_to_gpu(local)
info = {}
try:
    p = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    info["bus"] = bus
    info["numa_node_file"] = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
    info["nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
except Exception as e:
    info["error"] = repr(e)
info["affinity_cpus"] = len(os.sched_getaffinity(0))
UP, DOWN = 9568256, 28049408
h_up = torch.empty(UP // 4, dtype=torch.float32).pin_memory(); d_up = torch.empty(UP // 4, device=dev)
h_dn = torch.empty(DOWN // 4, dtype=torch.float32).pin_memory(); d_dn = torch.empty(DOWN // 4, device=dev)
s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def run(up, down, n=100):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if up:
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize()
    el = shard.max_over_ranks(time.perf_counter() - t0, dev)
    return el / n


for _ in range(2):
    run(True, True, 10)
t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    print(json.dumps({"world": world, "bind": bind, "topology": info,
                      "h2d_alone_gbs": UP / t_up / 1e9, "d2h_alone_gbs": DOWN / t_dn / 1e9,
                      "both_ms_per_step": t_both * 1e3, "both_h2d_gbs": UP / t_both / 1e9, "both_d2h_gbs": DOWN / t_both / 1e9,
                      "points_per_s_bound_per_gpu": 32768 / t_both}))
if world > 1:
    torch.distributed.destroy_process_group()
