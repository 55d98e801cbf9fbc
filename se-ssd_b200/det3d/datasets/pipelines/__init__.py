from .formating import Reformat
from .loading import LoadPointCloudAnnotations, LoadPointCloudFromFile
from .preprocess import AssignTarget, Voxelization

__all__ = ["Voxelization", "AssignTarget", "Reformat", "LoadPointCloudFromFile", "LoadPointCloudAnnotations"]
