"""avg_voxelize / spherical_avg_voxelize — same contracts as
/root/reference/PVCNN/modules/functional/voxelization.py:10-44 and spherical_vox.py:10-44."""
from torch.autograd import Function

from ..backend import _backend

__all__ = ['avg_voxelize', 'spherical_avg_voxelize']


class _AvgVoxelizeBase(Function):
    _fwd = None
    _bwd = None

    @classmethod
    def _run(cls, ctx, features, coords, resolution):
        b, c, n = features.shape
        out, indices, counts = cls._fwd(features, coords, resolution)
        ctx.mark_non_differentiable(indices)
        ctx.save_for_backward(indices, counts)
        return out.view(b, c, resolution, resolution, resolution), indices.view(b, n)

    @classmethod
    def _grad(cls, ctx, grad_output):
        b, c = grad_output.shape[:2]
        indices, counts = ctx.saved_tensors
        return cls._bwd(grad_output.contiguous().view(b, c, -1), indices, counts)


class AvgVoxelization(_AvgVoxelizeBase):
    """(features [B,C,N], int voxel coords [B,3,N], r) -> (voxel means [B,C,r,r,r], voxel index per point [B,N])."""
    _fwd = staticmethod(_backend.avg_voxelize_forward)
    _bwd = staticmethod(_backend.avg_voxelize_backward)

    @staticmethod
    def forward(ctx, features, coords, resolution):
        return AvgVoxelization._run(ctx, features.contiguous(), coords.int().contiguous(), resolution)

    @staticmethod
    def backward(ctx, grad_output, _g_ind):
        return AvgVoxelization._grad(ctx, grad_output), None, None


class SphericalAvgVoxelization(_AvgVoxelizeBase):
    """(features [B,C,N], normalised fp32 coords [B,3,N], r) -> (means on the (gamma,alpha,beta) grid, cell per point;
    -1 marks points outside the unit ball / on the pole)."""
    _fwd = staticmethod(_backend.spherical_avg_voxelize_forward)
    _bwd = staticmethod(_backend.spherical_avg_voxelize_backward)

    @staticmethod
    def forward(ctx, features, coords, resolution):
        return SphericalAvgVoxelization._run(ctx, features.contiguous(), coords.contiguous(), resolution)

    @staticmethod
    def backward(ctx, grad_output, _g_ind):
        return SphericalAvgVoxelization._grad(ctx, grad_output), None, None


avg_voxelize = AvgVoxelization.apply
spherical_avg_voxelize = SphericalAvgVoxelization.apply
