"""Voxel readers: VoxelFeatureExtractorV3 (the per-voxel mean; fused into the voxeliser on the engine path)."""
from .voxel_encoder import VoxelFeatureExtractorV3

__all__ = ["VoxelFeatureExtractorV3"]
