from .voxel_encoder import VoxelFeatureExtractorV3

__all__ = ["VoxelFeatureExtractorV3"]
