"""Barycentre grid subsampling on the GPU — the reference's `utils/grid_subsampleing.py:3-21` API over csrc/gridsub.cu
(reference implementation: a CPU CPython extension, cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp).

`grid_sub_sampling(points, features=None, labels=None, grid_size=0.1, verbose=0)` keeps the reference's signature and its
return convention (points; points, features; points, labels; or all three — `wrapper.cpp:265-284`).  numpy inputs come back
as numpy arrays (the reference's types), CUDA tensors come back as CUDA tensors.  Cells are returned in ascending cell
index (the reference: unordered_map order), barycentres and mean features are bit-identical to the reference's.
"""
import numpy as np
import torch


def _prepare(points, features, labels, dev):
    """Device tensors of one cloud in the shapes the op takes, with the reference wrapper's shape errors."""
    pts = torch.as_tensor(points, dtype=torch.float32).to(dev).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")                      # wrapper.cpp:118-125
    N = pts.shape[0]
    if features is None:
        f = torch.empty((N, 0), dtype=torch.float32, device=dev)
    else:
        f = torch.as_tensor(features, dtype=torch.float32).to(dev).contiguous()
        if f.dim() != 2 or f.shape[0] != N:
            raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")                # wrapper.cpp:127-145
    if labels is None:
        l = torch.empty((N, 0), dtype=torch.int32, device=dev)
    else:
        l = torch.as_tensor(labels).to(torch.int32).to(dev).contiguous()
        if l.dim() > 2 or l.shape[0] != N:
            raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")         # wrapper.cpp:147-162
        l = l.reshape(N, -1)
    return pts, f, l


def grid_sub_sampling_many(clouds, grid_size=0.1, device=None):
    """Several scans in one go: `clouds` is a list of `points` or `(points, features)` or `(points, features, labels)` (CUDA
    tensors or numpy arrays; features / labels may be None).  Every scan's kernels are enqueued back to back and the output
    lengths are read with ONE host synchronisation at the end (the single-cloud call reads its length after every scan, which
    leaves the GPU idle while the host prepares the next one).  Returns a list of tuples (points, features, labels) of CUDA
    tensors, None where the input had none; results identical to grid_sub_sampling()."""
    items = [c if isinstance(c, (tuple, list)) else (c,) for c in clouds]
    items = [tuple(c) + (None,) * (3 - len(c)) for c in items]
    if not items:
        return []
    first = items[0][0]
    dev = torch.device(device) if device is not None else (first.device if isinstance(first, torch.Tensor) else torch.device("cuda"))
    if dev.type != "cuda":
        raise RuntimeError("grid_sub_sampling runs on a CUDA device only (no CPU path exists in this package)")
    raw = []
    for (p, f, l) in items:
        pts, ft, lb = _prepare(p, f, l, dev)
        raw.append(torch.ops.ri.grid_subsample(pts, ft, lb, float(grid_size)) + (f is not None, l is not None))
    counts = torch.cat([r[3] for r in raw]).cpu().tolist()        # the only host synchronisation
    return [(op[:M], of[:M] if has_f else None, ol[:M] if has_l else None) for (op, of, ol, _, has_f, has_l), M in zip(raw, counts)]


def grid_sub_sampling(points, features=None, labels=None, grid_size=0.1, verbose=0, device=None):
    as_numpy = not isinstance(points, torch.Tensor)
    dev = torch.device(device) if device is not None else (torch.device("cuda") if as_numpy else points.device)
    if dev.type != "cuda":
        raise RuntimeError("grid_sub_sampling runs on a CUDA device only (no CPU path exists in this package)")
    pts, f, l = _prepare(points, features, labels, dev)
    op, of, ol, cnt = torch.ops.ri.grid_subsample(pts, f, l, float(grid_size))
    M = int(cnt.item())                                       # the only host synchronisation: the output length
    out = [op[:M]]
    if features is not None:
        out.append(of[:M])
    if labels is not None:
        out.append(ol[:M])
    if as_numpy:
        out = [o.cpu().numpy() for o in out]
    return out[0] if len(out) == 1 else tuple(out)
