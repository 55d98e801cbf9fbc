"""TargetAssigner (reference: det3d/core/anchor/target_assigner.py:8-181; the assign_v2 path the SE-SSD config uses)."""
from collections import OrderedDict

import numpy as np

from .target_ops_v2 import create_target_np


class TargetAssigner:
    def __init__(self, box_coder, anchor_generators, region_similarity_calculator=None, positive_fraction=None, sample_size=512):
        self._region_similarity_calculator = region_similarity_calculator
        self._box_coder = box_coder
        self._anchor_generators = anchor_generators
        self._positive_fraction = positive_fraction
        self._sample_size = sample_size

    box_coder = property(lambda self: self._box_coder)

    @property
    def classes(self):
        return [a.class_name for a in self._anchor_generators]

    def generate_anchors_dict(self, feature_map_size):
        out = OrderedDict()
        for a in self._anchor_generators:
            anchors = a.generate(feature_map_size)
            anchors = anchors.reshape([*anchors.shape[:3], -1, anchors.shape[-1]])
            num = int(np.prod(anchors.shape[:-1]))
            out[a.class_name] = {"anchors": anchors,
                                 "matched_thresholds": np.full(num, a.match_threshold, anchors.dtype),
                                 "unmatched_thresholds": np.full(num, a.unmatch_threshold, anchors.dtype)}
        return out

    def assign_v2(self, anchors_dict, gt_boxes, anchors_mask=None, gt_classes=None, gt_names=None, enable_similar_type=False):
        if anchors_mask is not None:
            raise NotImplementedError("anchors_mask is unused by the SE-SSD config")
        code = self._box_coder.code_size

        def similarity_fn(anchors, gts):
            return self._region_similarity_calculator.compare(anchors[:, [0, 1, 3, 4, -1]], gts[:, [0, 1, 3, 4, -1]])

        def box_encoding_fn(boxes, anchors):
            return self._box_coder.encode(boxes, anchors)

        results = []
        fmap = None
        for class_name, ad in anchors_dict.items():
            mask = np.array([c == class_name for c in gt_names], dtype=np.bool_)
            if enable_similar_type:                       # all GT classes collapse to label 1 (config.py:107)
                mask = np.ones(gt_names.shape, dtype=np.bool_)
                gt_classes = np.ones(gt_names.shape, dtype=np.int32)
            fmap = ad["anchors"].shape[:3]
            results.append(create_target_np(ad["anchors"].reshape(-1, code), gt_boxes[mask], similarity_fn, box_encoding_fn,
                                            gt_classes=gt_classes[mask], matched_threshold=ad["matched_thresholds"],
                                            unmatched_threshold=ad["unmatched_thresholds"],
                                            positive_fraction=self._positive_fraction, rpn_batch_size=self._sample_size,
                                            norm_by_num_examples=False, box_code_size=code))
        out = {"positive_gt_id": [t["positive_gt_id"] for t in results]}
        out["bbox_targets"] = np.concatenate([t["bbox_targets"].reshape(*fmap, -1, code) for t in results], axis=-2).reshape(-1, code)
        out["labels"] = np.concatenate([t["labels"].reshape(*fmap, -1) for t in results], axis=-1).reshape(-1)
        out["bbox_outside_weights"] = np.concatenate([t["bbox_outside_weights"].reshape(*fmap, -1) for t in results], axis=-1).reshape(-1)
        return out
