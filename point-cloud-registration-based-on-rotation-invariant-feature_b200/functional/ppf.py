"""ppf — same contract as /root/reference/PVCNN/modules/functional/ppf.py:8-22 (plain function, no autograd)."""
import torch

from ..backend import _backend

__all__ = ['ppf', 'knn_ppf', 'knn_ppf_fused']


def ppf(centers_coords, points_coords, centers_normals, points_normals):
    """All four FloatTensor[B,3,L] -> FloatTensor[B,4,L]:
    (angle(d, n_centre), angle(d, n_point), angle(n_centre, n_point), ||d||) with d = centre - point.
    The backend takes (points, centres, point normals, centre normals) — the swap of reference ppf.py:22."""
    return _backend.spherical_ppf_forward(points_coords.contiguous(), centers_coords.contiguous(),
                                          points_normals.contiguous(), centers_normals.contiguous())


def knn_ppf(xyz, normals, idx):
    """Fused neighbour gather + PPF: xyz, normals [B,3,N], idx int32 [B,k,N] -> [B,4,k,N].
    Identical values to ppf(xyz.expand_k, gather(xyz, idx), normals.expand_k, gather(normals, idx)) reshaped."""
    return torch.ops.ri.ppf_gather(xyz.contiguous(), normals.contiguous(), idx.contiguous())


def knn_ppf_fused(xyz, normals, k):
    """k nearest neighbours of every point inside its own cloud and the PPF of every (point, neighbour) pair, one kernel:
    xyz, normals [B,3,N] -> (dist [B,k,N], idx [B,k,N], ppf [B,4,k,N]); equals k_nearest_neighbor + knn_ppf."""
    return torch.ops.ri.knn_ppf(xyz.float().contiguous(), normals.float().contiguous(), int(k))
