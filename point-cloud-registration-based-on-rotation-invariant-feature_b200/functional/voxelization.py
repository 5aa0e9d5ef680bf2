"""avg_voxelize / spherical_avg_voxelize — same contracts as
/root/reference/PVCNN/modules/functional/voxelization.py:10-44 and spherical_vox.py:10-44 — plus the fused
voxelize + DGCNN-edge forms this package adds."""
import torch
from torch.autograd import Function

from ..backend import _backend

__all__ = ['avg_voxelize', 'spherical_avg_voxelize', 'avg_voxelize_edge', 'spherical_avg_voxelize_edge']


def _forward(ctx, fwd, features, coords, resolution):
    b, c, n = features.shape
    out, indices, counts = fwd(features, coords, resolution)
    ctx.mark_non_differentiable(indices)
    ctx.save_for_backward(indices, counts)
    return out.view(b, c, resolution, resolution, resolution), indices.view(b, n)


def _backward(ctx, bwd, grad_output):
    b, c = grad_output.shape[:2]
    indices, counts = ctx.saved_tensors
    return bwd(grad_output.contiguous().view(b, c, -1), indices, counts)


class AvgVoxelization(Function):
    """(features [B,C,N], int voxel coords [B,3,N], r) -> (voxel means [B,C,r,r,r], voxel index per point [B,N])."""

    @staticmethod
    def forward(ctx, features, coords, resolution):
        return _forward(ctx, _backend.avg_voxelize_forward, features.contiguous(), coords.int().contiguous(), resolution)

    @staticmethod
    def backward(ctx, grad_output, _g_ind):
        return _backward(ctx, _backend.avg_voxelize_backward, grad_output), None, None


class SphericalAvgVoxelization(Function):
    """(features [B,C,N], normalised fp32 coords [B,3,N], r) -> (means on the (gamma,alpha,beta) grid, cell per point;
    -1 marks points outside the unit ball / on the pole)."""

    @staticmethod
    def forward(ctx, features, coords, resolution):
        return _forward(ctx, _backend.spherical_avg_voxelize_forward, features.contiguous(), coords.contiguous(), resolution)

    @staticmethod
    def backward(ctx, grad_output, _g_ind):
        return _backward(ctx, _backend.spherical_avg_voxelize_backward, grad_output), None, None


# ---- fused voxelize + DGCNN edge features ---------------------------------------------------------------
def _forward_edge(ctx, op, features, coords, resolution):
    b, c, n = features.shape
    out, indices, counts, edge = op(features, coords, resolution)
    ctx.mark_non_differentiable(indices)
    ctx.save_for_backward(indices, counts)
    return out.view(b, c, resolution, resolution, resolution), indices.view(b, n), edge


def _backward_edge(ctx, g_grid, g_edge):
    """edge = cat(features - grid[:, :, ind] (0 where ind == -1), features): route the edge gradient back into the
    grid gradient, then one voxelize backward."""
    indices, counts = ctx.saved_tensors
    b, c2, n = g_edge.shape
    c = c2 // 2
    g_rel = g_edge[:, :c, :].masked_fill((indices == -1).unsqueeze(1), 0.0)
    g_avg = g_grid.contiguous().view(b, c, -1).clone()
    g_avg.scatter_add_(2, indices.clamp(min=0).long().unsqueeze(1).expand(-1, c, -1), -g_rel)
    return torch.ops.ri.voxelize_backward(g_avg, indices, counts) + g_rel + g_edge[:, c:, :]


class AvgVoxelizationEdge(Function):
    """(features, int voxel coords, r) -> (grid [B,C,r,r,r], ind [B,N], edge [B,2C,N]) in one pass;
    edge == voxel_edge_features(grid, features, ind)  (reference: voxelize, then pvconv.py:68-90)."""

    @staticmethod
    def forward(ctx, features, coords, resolution):
        return _forward_edge(ctx, torch.ops.ri.cube_voxelize_edge, features.contiguous(), coords.int().contiguous(), resolution)

    @staticmethod
    def backward(ctx, g_grid, _g_ind, g_edge):
        return _backward_edge(ctx, g_grid, g_edge), None, None


class SphericalAvgVoxelizationEdge(Function):
    @staticmethod
    def forward(ctx, features, coords, resolution):
        return _forward_edge(ctx, torch.ops.ri.sph_voxelize_edge, features.contiguous(), coords.contiguous(), resolution)

    @staticmethod
    def backward(ctx, g_grid, _g_ind, g_edge):
        return _backward_edge(ctx, g_grid, g_edge), None, None


avg_voxelize = AvgVoxelization.apply
spherical_avg_voxelize = SphericalAvgVoxelization.apply
avg_voxelize_edge = AvgVoxelizationEdge.apply
spherical_avg_voxelize_edge = SphericalAvgVoxelizationEdge.apply
