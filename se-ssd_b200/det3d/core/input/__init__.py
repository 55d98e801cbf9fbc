from .voxel_generator import VoxelGenerator

__all__ = ["VoxelGenerator"]
