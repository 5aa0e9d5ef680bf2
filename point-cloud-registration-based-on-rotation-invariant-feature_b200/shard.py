"""Multi-GPU sharding of the front end: one process per GPU, clouds (or source/target pairs) split by rank,
results gathered once at the end.  Replaces the reference's single-process nn.DataParallel
(/root/reference/train.py:116).  Every kernel on the path is independent per cloud (blockIdx = cloud index in all
reference kernels, e.g. knn.cu:9), so there is NO data-path collective: NCCL is used only by `gather_clouds`.
The same functions run over gloo on CPU tensors (used by the world_size-2 tests)."""
import os

import torch
import torch.distributed as dist

__all__ = ['init_from_env', 'shard_range', 'gather_clouds', 'max_over_ranks', 'bind_host_to_gpu']


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local_rank).
    Without WORLD_SIZE in the environment this is a single-process run and nothing is initialised."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) of `total` clouds for `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_clouds(local, total, dim=0):
    """All-gather per-rank result slabs (split along `dim` by shard_range) back into the full batch on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    if dim != 0:
        local = local.transpose(0, dim)
    local = local.contiguous()
    pad = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)
    return full.transpose(0, dim) if dim != 0 else full


def max_over_ranks(value, device):
    """Max of a python float over all ranks (device timing rule: report the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(index):
    """Host side of a rank that feeds GPU `index` over PCIe: restrict the process to the CPU cores of the NUMA node the GPU
    hangs off and prefer that node for new pages, BEFORE any pinned buffer is allocated — eight ranks that stage through
    the memory of one socket share its controllers and cross the socket link for half of the GPUs.  Returns
    {'node', 'cpus', 'mempolicy'} or None when the topology cannot be read (containers without /sys, single-node hosts)."""
    try:
        props = torch.cuda.get_device_properties(index)
        bus = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open('/sys/devices/system/node/node%d/cpulist' % node).read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        policy = False
        try:                                            # set_mempolicy(MPOL_PREFERRED, {node}): syscall 238 on x86-64
            import ctypes
            import platform
            if platform.machine() == 'x86_64' and node < 64:
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << node)
                policy = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(65)) == 0
        except Exception:
            policy = False
        return {'node': node, 'cpus': len(cpus), 'mempolicy': bool(policy)}
    except Exception:
        return None
