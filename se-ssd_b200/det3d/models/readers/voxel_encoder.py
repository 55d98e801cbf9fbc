"""VoxelFeatureExtractorV3: per-voxel mean of the (<= max_points) points (reference: det3d/models/readers/voxel_encoder.py:197-210).
Module-level API; on the fused device path the mean is produced by the voxeliser kernel itself (sessd_voxelize, d_mean)."""
from torch import nn

from ..registry import READERS


@READERS.register_module
class VoxelFeatureExtractorV3(nn.Module):
    def __init__(self, num_input_features=4, norm_cfg=None, name="VoxelFeatureExtractorV3"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, voxels, num_points_per_voxel, coors=None):
        total = voxels[:, :, : self.num_input_features].sum(dim=1, keepdim=False)
        return (total / num_points_per_voxel.type_as(voxels).view(-1, 1)).contiguous()
