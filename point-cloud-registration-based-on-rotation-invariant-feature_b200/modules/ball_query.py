"""BallQuery — neighbourhood grouping module with the reference's contract (PVCNN/modules/ball_query.py:10-35 of the reference):
constructor `(radius, num_neighbors, include_coordinates=True)`, `forward(points_coords, centers_coords, points_features=None)`
returning the grouped tensor laid out `[B, C', U, M]` (neighbours before centres), where C' is 3 relative coordinates and / or the
grouped feature channels.  Everything heavy runs in csrc/ballquery.cu (`ri_ball_query_f32`, `ri_grouping_f32`).

For the models' local PPF branch prefer `functional.ball_local_ppf`, which never materialises the grouped tensors."""
import torch
from torch import nn

from .. import functional as F

__all__ = ['BallQuery']


class BallQuery(nn.Module):
    def __init__(self, radius, num_neighbors, include_coordinates=True):
        super().__init__()
        self.radius, self.num_neighbors, self.include_coordinates = radius, num_neighbors, include_coordinates

    def neighbours(self, points_coords, centers_coords):
        """int32 [B, M, U]: the first `num_neighbors` points inside the ball of each centre, in index order (self excluded)."""
        return F.ball_query(centers_coords.contiguous(), points_coords.contiguous(), self.radius, self.num_neighbors)

    def forward(self, points_coords, centers_coords, points_features=None):
        if points_features is None and not self.include_coordinates:
            raise AssertionError('No Features For Grouping')
        idx = self.neighbours(points_coords, centers_coords)
        channels = []
        if self.include_coordinates or points_features is None:
            # coordinates of the neighbours relative to their centre, [B, 3, M, U]
            channels.append(F.grouping(points_coords.contiguous(), idx) - centers_coords.contiguous()[..., None])
        if points_features is not None:
            channels.append(F.grouping(points_features, idx))                 # [B, C, M, U]
        grouped = channels[0] if len(channels) == 1 else torch.cat(channels, dim=1)
        return grouped.transpose(2, 3)                                        # [B, C', U, M]

    def extra_repr(self):
        tail = ', include coordinates' if self.include_coordinates else ''
        return f'radius={self.radius}, num_neighbors={self.num_neighbors}{tail}'
