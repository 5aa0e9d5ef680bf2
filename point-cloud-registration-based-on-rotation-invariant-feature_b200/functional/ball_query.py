"""ball_query / grouping — same contracts as /root/reference/PVCNN/modules/functional/ball_query.py:8-19 and
grouping.py:8-33."""
import torch

from ..backend import _backend

__all__ = ['ball_query', 'grouping']


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """centers_coords FloatTensor[B,3,M], points_coords FloatTensor[B,3,N] -> neighbour indices IntTensor[B,M,U]:
    the first U points in index order with 1e-5 < d^2 < radius^2 (first hit replicated over the row, zeros if none)."""
    return _backend.ball_query(centers_coords.contiguous(), points_coords.contiguous(), radius, num_neighbors)


class Grouping(torch.autograd.Function):
    """features FloatTensor[B,C,N], indices IntTensor[B,M,U] -> grouped FloatTensor[B,C,M,U]."""

    @staticmethod
    def forward(ctx, features, indices):
        features = features.contiguous()
        indices = indices.contiguous()
        ctx.save_for_backward(indices)
        ctx.num_points = features.size(-1)
        return _backend.grouping_forward(features, indices)

    @staticmethod
    def backward(ctx, grad_output):
        indices, = ctx.saved_tensors
        return _backend.grouping_backward(grad_output.contiguous(), indices, ctx.num_points), None


grouping = Grouping.apply
