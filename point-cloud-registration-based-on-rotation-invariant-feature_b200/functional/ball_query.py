"""ball_query / grouping — same contracts as /root/reference/PVCNN/modules/functional/ball_query.py:8-19 and
grouping.py:8-33."""
import torch

from ..backend import _backend

__all__ = ['ball_query', 'grouping', 'local_ppf', 'ball_local_ppf', 'fold_fuser', 'local_ppf_features']


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


def fold_fuser(fuser):
    """BatchNorm-folded weights of the models' local-feature MLP `fuser = SharedMLP(4, [32, 64], dim=2)`
    (/root/reference/PVCNN/models/pvcnn_classify.py:65-66; modules/shared_mlp.py:6-31: Conv2d(1x1) BN ReLU, twice) for eval-mode
    inference: y = relu(W' x + b') with W' = W * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(var + eps) + beta.
    Accepts the reference's module or this package's SharedMLP (same `layers` Sequential).  -> (w1 [32,4], b1, w2 [64,32], b2)."""
    layers = list(fuser.layers)
    convs = [m for m in layers if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d))]
    bns = [m for m in layers if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d))]
    if len(convs) != 2 or len(bns) != 2:
        raise ValueError("expected SharedMLP(in, [c1, c2]): two 1x1 convolutions, each followed by a BatchNorm")
    out = []
    with torch.no_grad():
        for conv, bn in zip(convs, bns):
            w = conv.weight.reshape(conv.out_channels, conv.in_channels).double()
            b = (conv.bias if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)).double()
            g = (bn.weight.double() if bn.affine else torch.ones_like(b)) / torch.sqrt(bn.running_var.double() + bn.eps)
            beta = bn.bias.double() if bn.affine else torch.zeros_like(b)
            out += [(w * g[:, None]).float().contiguous(), ((b - bn.running_mean.double()) * g + beta).float().contiguous()]
    return tuple(out)


def local_ppf_features(coords, normals, folded, radius=0.3, num_neighbors=128, centers_coords=None, centers_normals=None):
    """`self.fuser(local_ppf).max(dim=2).values` of the shipped models (pvcnn_classify.py:252-271) in eval mode, from raw
    coordinates and normals: ball query, then ONE kernel for point-pair features -> MLP -> max over the neighbours
    (csrc/localmlp.cu).  `folded` = fold_fuser(model.fuser).  -> FloatTensor[B, 64, M]."""
    cc = coords if centers_coords is None else centers_coords
    cn = normals if centers_normals is None else centers_normals
    idx = ball_query(cc, coords, radius, num_neighbors)
    w1, b1, w2, b2 = folded
    return torch.ops.ri.local_ppf_mlp_max(coords.float().contiguous(), normals.float().contiguous(), cc.float().contiguous(),
                                          cn.float().contiguous(), idx, w1, b1, w2, b2)
