"""voxel_edge_features — the DGCNN voxel-neighbourhood grouping of PVConv.forward
(/root/reference/PVCNN/modules/pvconv.py:68-90) as ONE op instead of ~8 torch kernels, a deepcopy and two host syncs."""
import torch
from torch.autograd import Function

__all__ = ['voxel_edge_features']


class VoxelEdgeFeatures(Function):
    """(avg_voxel_features [B,C,r,r,r] or [B,C,s], features [B,C,N], inds [B,N]) ->
    cat(features - avg[:, :, inds] (zeros where inds == -1), features)  [B,2C,N]"""

    @staticmethod
    def forward(ctx, avg_voxel_features, features, inds):
        B, C, N = features.shape
        avg = avg_voxel_features.contiguous().view(B, C, -1)
        out = torch.ops.ri.voxel_edge_gather(avg, features.contiguous(), inds.contiguous())
        ctx.save_for_backward(inds)
        ctx.grid_shape = avg_voxel_features.shape
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (inds,) = ctx.saved_tensors
        B, C2, N = grad_out.shape
        C = C2 // 2
        mask = (inds == -1).unsqueeze(1)
        g_rel = grad_out[:, :C, :].masked_fill(mask, 0.0)
        g_feat = g_rel + grad_out[:, C:, :]
        s = 1
        for d in ctx.grid_shape[2:]:
            s *= d
        g_avg = torch.zeros((B, C, s), dtype=grad_out.dtype, device=grad_out.device)
        index = inds.clamp(min=0).long().unsqueeze(1).expand(-1, C, -1)
        g_avg.scatter_add_(2, index, -g_rel)
        return g_avg.view(ctx.grid_shape), g_feat, None


voxel_edge_features = VoxelEdgeFeatures.apply
