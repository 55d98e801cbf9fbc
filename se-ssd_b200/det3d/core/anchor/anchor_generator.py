"""Anchor generators (reference: det3d/core/anchor/anchor_generator.py:64-117)."""
import numpy as np

from det3d.core.bbox import box_np_ops


class AnchorGeneratorRange:
    def __init__(self, anchor_ranges, sizes=[1.6, 3.9, 1.56], rotations=[0, np.pi / 2], velocities=None, class_name=None,
                 match_threshold=-1, unmatch_threshold=-1, dtype=np.float32):
        self._sizes, self._anchor_ranges, self._rotations = sizes, anchor_ranges, rotations
        self._velocities, self._dtype, self._class_name = velocities, dtype, class_name
        self._match_threshold, self._unmatch_threshold = match_threshold, unmatch_threshold
        self._anchors = None

    class_name = property(lambda self: self._class_name)
    match_threshold = property(lambda self: self._match_threshold)
    unmatch_threshold = property(lambda self: self._unmatch_threshold)

    @property
    def num_anchors_per_localization(self):
        return len(self._rotations) * np.array(self._sizes).reshape([-1, 3]).shape[0]

    @property
    def ndim(self):
        return self._anchors.shape[-1]

    def generate(self, feature_map_size):
        self._anchors = box_np_ops.create_anchors_3d_range(feature_map_size, self._anchor_ranges, self._sizes, self._rotations,
                                                           self._velocities, self._dtype)
        return self._anchors
