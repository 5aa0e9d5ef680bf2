"""k_nearest_neighbor — same contract as /root/reference/PVCNN/modules/functional/knn.py:8-27."""
import torch

from ..backend import _backend

__all__ = ['k_nearest_neighbor', 'knn_indices']


class KNearestNeighbor(torch.autograd.Function):
    """(xyz1 [B,c,n], xyz2 [B,c,m], k) -> (dist1 [B,k,n], dist2 [B,k,m], idx1, idx2); squared L2 distances ascending
    along k, int32 indices (non-differentiable); gradients flow to both point sets through the distances."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, k):
        xyz1 = xyz1.float().contiguous()
        xyz2 = xyz2.float().contiguous()
        dist1, dist2, idx1, idx2 = _backend.knn_forward_cuda(xyz1, xyz2, k)
        ctx.mark_non_differentiable(idx1, idx2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, _g_idx1, _g_idx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1, gradxyz2 = _backend.knn_backward_cuda(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(),
                                                        idx1, idx2)
        return gradxyz1, gradxyz2, None


k_nearest_neighbor = KNearestNeighbor.apply


def knn_indices(xyz, k):
    """Self-query convenience used by the fused front end: (dist [B,k,N], idx [B,k,N]) of xyz against itself,
    one direction only (the bilateral op would compute the same thing twice)."""
    return torch.ops.ri.knn_one(xyz.float().contiguous(), xyz.float().contiguous(), int(k))
