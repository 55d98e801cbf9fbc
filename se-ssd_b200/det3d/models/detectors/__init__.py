from .single_stage import SingleStageDetector
from .voxelnet_sessd import VoxelNet

__all__ = ["SingleStageDetector", "VoxelNet"]
