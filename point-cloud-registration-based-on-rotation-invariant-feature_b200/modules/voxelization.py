"""Voxelization / Spherical_Voxelization — nn.Module shells with the contracts of
/root/reference/PVCNN/modules/voxelization.py:9-38 and spherical_vox.py:9-26.

The coordinate prologue (centre, scale, clamp, round) stays in torch on purpose: the voxel indices are only
bit-exact against the reference if the normalised coordinates fed to the binning kernel are bit-identical, and
torch's own mean/norm/max reductions are the definition of those bits."""
import torch
import torch.nn as nn

from .. import functional as F

__all__ = ['Voxelization', 'Spherical_Voxelization']


class Voxelization(nn.Module):
    def __init__(self, resolution, normalize=True, eps=0):
        super().__init__()
        self.r = int(resolution)
        self.normalize = normalize
        self.eps = eps

    def forward(self, features, coords, with_edge=False):
        """features [B,C,N], coords [B,3,N] -> (voxel means [B,C,r,r,r], voxel index per point [B,N],
        continuous grid coordinates [B,3,N] in [0, r-1] (what trilinear_devoxelize consumes)).
        with_edge=True (extension) appends the DGCNN edge features [B,2C,N] produced in the same pass."""
        coords = coords.detach()
        norm_coords = coords - coords.mean(2, keepdim=True)
        if self.normalize:
            radius = norm_coords.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values
            norm_coords = norm_coords / (radius * 2.0 + self.eps) + 0.5
        else:
            norm_coords = (norm_coords + 1) / 2.0
        norm_coords = torch.clamp(norm_coords * self.r, 0, self.r - 1)
        vox_coords = torch.round(norm_coords).to(torch.int32)
        if with_edge:
            out, indices, edge = F.avg_voxelize_edge(features, vox_coords, self.r)
            return out, indices.detach(), norm_coords, edge
        out, indices = F.avg_voxelize(features, vox_coords, self.r)
        return out, indices.detach(), norm_coords

    def extra_repr(self):
        return 'resolution={}{}'.format(self.r, ', normalized eps = {}'.format(self.eps) if self.normalize else '')


class Spherical_Voxelization(nn.Module):
    def __init__(self, resolution):
        super().__init__()
        self.r = int(resolution)

    def forward(self, features, coords, with_edge=False):
        """features [B,C,N], coords [B,3,N] -> (means on the spherical grid [B,C,r,r,r], cell per point [B,N]
        (-1 = undefined), centred coords scaled so the farthest point has radius ~1).
        with_edge=True (extension) appends the DGCNN edge features [B,2C,N] produced in the same pass."""
        coords = coords.detach()
        norm_coords = coords - coords.mean(2, keepdim=True)
        norm_coords = norm_coords / (norm_coords.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values + 1e-20)
        if with_edge:
            out, inds, edge = F.spherical_avg_voxelize_edge(features, norm_coords, self.r)
            return out, inds.detach(), norm_coords, edge
        out, inds = F.spherical_avg_voxelize(features, norm_coords, self.r)
        return out, inds.detach(), norm_coords

    def extra_repr(self):
        return 'resolution={}'.format(self.r)
