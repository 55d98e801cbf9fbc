"""``points_to_voxel`` with the reference's functional signature (det3d/ops/point_cloud/point_cloud_ops_v2.py:120-194)."""
import numpy as np

from sessd_b200 import ops


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000):
    if not reverse_index:
        raise NotImplementedError("only the zyx (reverse_index=True) layout used by VoxelGenerator is provided")
    points = np.ascontiguousarray(points, dtype=np.float32)
    cfg = ops.make_voxel_cfg(voxel_size, coors_range, max_points, max_voxels, points.shape[1])
    return ops.voxelize_host(points, cfg)
