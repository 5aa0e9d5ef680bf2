"""ball_query / grouping — same contracts as /root/reference/PVCNN/modules/functional/ball_query.py:8-19 and
grouping.py:8-33."""
import torch

from ..backend import _backend

__all__ = ['ball_query', 'grouping', 'local_ppf', 'ball_local_ppf']


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


def local_ppf(points_coords, points_normals, neighbor_indices, centers_coords=None, centers_normals=None):
    """The `with_local_feat == 'ppf'` features of the shipped models (/root/reference/PVCNN/models/pvcnn_classify.py:252-270)
    straight from the ball-query indices, one kernel, nothing grouped in memory: FloatTensor[B,4,U,M] =
    (angle(n_nbr, d), angle(n_centre, d), angle(n_nbr, n_centre), |d|) with the reference's d = c - (p_nbr - c).
    Centres default to the points themselves (what the models do: `center_coords = coords`)."""
    cc = points_coords if centers_coords is None else centers_coords
    cn = points_normals if centers_normals is None else centers_normals
    return torch.ops.ri.local_ppf(points_coords.float().contiguous(), points_normals.float().contiguous(),
                                  cc.float().contiguous(), cn.float().contiguous(), neighbor_indices.contiguous())


def ball_local_ppf(coords, normals, radius=0.3, num_neighbors=128):
    """ball query (self excluded, first `num_neighbors` in index order) + local_ppf: what BallQuery + the torch PPF block compute
    for `neighbor_num=128, radius=0.3` (pvcnn_classify.py:61-63)."""
    idx = ball_query(coords, coords, radius, num_neighbors)
    return local_ppf(coords, normals, idx)
