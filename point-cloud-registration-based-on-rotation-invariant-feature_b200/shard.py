"""Multi-GPU sharding of the front end: one process per GPU, clouds (or source/target pairs) split by rank,
results gathered once at the end.  Replaces the reference's single-process nn.DataParallel
(/root/reference/train.py:116).  Every kernel on the path is independent per cloud (blockIdx = cloud index in all
reference kernels, e.g. knn.cu:9), so there is NO data-path collective: NCCL is used only by `gather_clouds`.
The same functions run over gloo on CPU tensors (used by the world_size-2 tests)."""
import os

import torch
import torch.distributed as dist

__all__ = ['init_from_env', 'shard_range', 'gather_clouds', 'max_over_ranks']


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
