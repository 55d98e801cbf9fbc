"""VoxelGenerator -- same constructor / ``generate`` / properties as det3d/core/input/voxel_generator.py:10-48, executed by the
CUDA voxeliser (sessd_voxelize_host) instead of the sequential numba loop of det3d/ops/point_cloud/point_cloud_ops_v2.py:9-62.
Outputs are bit-identical to the reference (tests/test_gpu_voxelize.py).  numpy in, fresh numpy out, exactly like the
reference; for the zero-copy batched device path use ``generate_batch``."""
import numpy as np

from sessd_b200 import ops


class VoxelGenerator:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        self._point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32).copy()
        self._voxel_size = np.asarray(voxel_size, dtype=np.float32).copy()
        extent = self._point_cloud_range[3:] - self._point_cloud_range[:3]
        self._grid_size = np.round(extent / self._voxel_size).astype(np.int64)          # (x, y, z) cells, :17-18
        self._max_num_points, self._max_voxels = max_num_points, max_voxels
        self._cfgs = {}                                                                  # one device config per point feature count

    def _cfg(self, num_feat):
        if num_feat not in self._cfgs:
            self._cfgs[num_feat] = ops.make_voxel_cfg(self._voxel_size, self._point_cloud_range, self._max_num_points,
                                                      self._max_voxels, num_feat)
        return self._cfgs[num_feat]

    def generate(self, points, max_voxels=20000):
        """points [N, >=3] float32 -> (voxels [M, max_points, F], coordinates [M, 3] zyx int32, num_points [M] int32).
        Like the reference (:24-32) the ``max_voxels`` argument is ignored in favour of the constructor's value."""
        points = np.ascontiguousarray(points, dtype=np.float32)
        return ops.voxelize_host(points, self._cfg(points.shape[1]))

    def generate_batch(self, points, frame_offsets, buffers=None):
        """Device path: points [P,F] cuda float32 (frames concatenated), frame_offsets [B+1] cuda int32.  Returns the
        capacity-sized ``ops.VoxelBuffers`` (voxels, coors with batch column, num_points, mean, num_voxels) -- the wire format of
        collate_kitti (det3d/torchie/parallel/collate.py:154-218) without leaving the GPU."""
        batch = int(frame_offsets.numel()) - 1
        if buffers is None:
            buffers = ops.VoxelBuffers(self._cfg(points.shape[1]), batch, max(int(points.shape[0]), 1), points.device)
        return ops.voxelize(points, frame_offsets, buffers)

    # read-only views of the constructor arguments (same names as the reference's properties)
    voxel_size = property(lambda self: self._voxel_size)
    max_num_points_per_voxel = property(lambda self: self._max_num_points)
    point_cloud_range = property(lambda self: self._point_cloud_range)
    grid_size = property(lambda self: self._grid_size)
