"""Module API — the hot-path classes of /root/reference/PVCNN/modules/__init__.py:1-10."""
from .pvconv import PVConv
from .shared_mlp import SharedMLP, SE3d
from .voxelization import Voxelization, Spherical_Voxelization
from .knn import knnModule
from .ball_query import BallQuery
