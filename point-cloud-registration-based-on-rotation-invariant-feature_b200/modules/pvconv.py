"""PVConv — contract of /root/reference/PVCNN/modules/pvconv.py:15-99 (same ctor arguments, same forward tuple,
same state_dict keys: voxel_layers.*, point_layers.*, coefficient), running on the ri_b200 ops.

Difference in mechanism only: the DGCNN voxel-neighbour grouping (pvconv.py:68-90: deepcopy + int64 index
expansion + gather + boolean-mask scatter + cat + two `.sum() > 0` host synchronisations) is emitted by the
voxelizer itself in the same pass (`with_edge=True`), so the dense grid is never gathered from.
`functional.voxel_edge_features` is the standalone form of the same computation."""
import torch
import torch.nn as nn

from .. import functional as F
from .voxelization import Voxelization, Spherical_Voxelization
from .shared_mlp import SharedMLP, SE3d

__all__ = ['PVConv']


class PVConv(nn.Module):
    def __init__(self, in_channels, out_channels, point_kernel_formal, voxel_shape, kernel_size, resolution,
                 with_coeff=False, with_se=False, normalize=True, eps=0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.point_kernel_formal = point_kernel_formal
        self.kernel_size = kernel_size
        self.resolution = resolution
        self.voxel_shape = voxel_shape
        self.with_coeff = with_coeff
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        self.spherical_vox = Spherical_Voxelization(resolution)
        if self.with_coeff:
            self.coefficient = nn.Parameter(torch.Tensor([1]))
        pad = kernel_size // 2
        voxel_layers = [
            nn.Conv3d(in_channels, out_channels, kernel_size, stride=1, padding=pad),
            nn.BatchNorm3d(out_channels, eps=1e-4),
            nn.LeakyReLU(0.1, True),
            nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=pad),
            nn.BatchNorm3d(out_channels, eps=1e-4),
            nn.LeakyReLU(0.1, True),
        ]
        if with_se:
            voxel_layers.append(SE3d(out_channels))
        self.voxel_layers = nn.Sequential(*voxel_layers)
        point_in = in_channels * 2 if point_kernel_formal == 'dgcnn_kernel' else in_channels
        self.point_layers = SharedMLP(point_in, out_channels)

    def forward(self, inputs):
        features, coords = inputs
        dgcnn = self.point_kernel_formal == 'dgcnn_kernel'
        edge = None
        if self.voxel_shape == 'cube':
            avg_voxel_features, inds, voxel_coords, *edge = self.voxelization(features, coords, with_edge=dgcnn)
            voxel_features = self.voxel_layers(avg_voxel_features)
            voxel_features = F.trilinear_devoxelize(voxel_features, voxel_coords, self.resolution, self.training)
        elif self.voxel_shape == 'spherical':
            avg_voxel_features, inds, voxel_coords, *edge = self.spherical_vox(features, coords, with_edge=dgcnn)
            voxel_features = self.voxel_layers(avg_voxel_features)
            voxel_features = F.spherical_trilinear_devoxelize(voxel_features, voxel_coords, inds, self.resolution,
                                                              self.training)
        else:
            raise ValueError('voxel_shape must be "cube" or "spherical"')
        if dgcnn:
            point_features = self.point_layers(edge[0])
        elif self.point_kernel_formal == 'pointnet_kernel':
            point_features = self.point_layers(features)
        else:
            raise ValueError('point_kernel_formal must be "dgcnn_kernel" or "pointnet_kernel"')
        if self.with_coeff:
            return self.coefficient * voxel_features + point_features, coords
        return voxel_features + point_features, coords
