from .collate import collate_kitti

__all__ = ["collate_kitti"]
