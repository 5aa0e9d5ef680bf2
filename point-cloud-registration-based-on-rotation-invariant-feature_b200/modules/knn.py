"""knnModule — contract of /root/reference/PVCNN/modules/knn.py:4-26."""
import torch.nn as nn

from ..functional.knn import k_nearest_neighbor

__all__ = ['knnModule']


class knnModule(nn.Module):
    def forward(self, input1, input2, k, bilateral, return_distance, return_index):
        """input1 [B,c,n], input2 [B,c,m].  Distances are returned as sqrt of the op's squared distances.
        (dist, idx) x (one direction | both) selected by the three flags; None if neither is requested."""
        dist1, dist2, idx1, idx2 = k_nearest_neighbor(input1, input2, k)
        picked = []
        if return_distance:
            picked += [dist1.sqrt(), dist2.sqrt()] if bilateral else [dist1.sqrt()]
        if return_index:
            picked += [idx1, idx2] if bilateral else [idx1]
        if not picked:
            return None
        return picked[0] if len(picked) == 1 else tuple(picked)
