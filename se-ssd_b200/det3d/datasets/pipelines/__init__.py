from .formating import Reformat
from .preprocess import AssignTarget, Voxelization

__all__ = ["Voxelization", "AssignTarget", "Reformat"]
