"""Region similarity calculators (reference: det3d/core/bbox/region_similarity.py:75-98; only the one the config uses)."""
from . import box_np_ops


class RegionSimilarityCalculator(object):
    def compare(self, boxes1, boxes2):
        return self._compare(boxes1, boxes2)


class NearestIouSimilarity(RegionSimilarityCalculator):
    """IoU of the nearest axis-aligned ("standing" or "lying") boxes of two rotated BEV box lists."""

    def _compare(self, boxes1, boxes2):
        return box_np_ops.iou_jit(box_np_ops.rbbox2d_to_near_bbox(boxes1), box_np_ops.rbbox2d_to_near_bbox(boxes2), eps=0.0)
