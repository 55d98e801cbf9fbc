"""Detectors named by the SE-SSD config: VoxelNet (reader -> sparse backbone -> neck -> head), built on SingleStageDetector."""
from .single_stage import SingleStageDetector
from .voxelnet_sessd import VoxelNet

__all__ = ["SingleStageDetector", "VoxelNet"]
