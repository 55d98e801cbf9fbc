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

    def assign_batch_gpu(self, anchors_dict, gt_boxes_list, device="cuda", max_gt=None, buffers=None):
        """Device twin of ``assign_v2`` for a whole batch of frames: single anchor class,
        ``enable_similar_type=True`` semantics (every GT box is class 1), thresholds taken from the anchor generator.
        ``gt_boxes_list``: one [M_i, 7] array per frame (already filtered / yaw-normalised like AssignTarget._assign).
        Returns device tensors ``labels [B,A] i32, bbox_targets [B,A,7], bbox_outside_weights [B,A]`` plus the compacted
        positives ``pos_anchor / positive_gt_id [B,A]`` with their per-frame count ``num_pos [B]`` (no host sync)."""
        import torch
        from sessd_b200 import ops
        if len(anchors_dict) != 1:
            raise NotImplementedError("assign_batch_gpu supports the single-class SE-SSD KITTI config")
        (ad,) = anchors_dict.values()
        gen = self._anchor_generators[0]
        anchors = np.ascontiguousarray(ad["anchors"].reshape(-1, self._box_coder.code_size), np.float32)
        B = len(gt_boxes_list)
        max_gt = int(max_gt or max([1] + [len(g) for g in gt_boxes_list]))
        gt = np.zeros((B, max_gt, 7), np.float32)
        num = np.zeros((B,), np.int32)
        for b, g in enumerate(gt_boxes_list):
            if len(g) > max_gt:
                raise ValueError("frame %d has %d GT boxes > max_gt=%d" % (b, len(g), max_gt))
            gt[b, :len(g)] = g
            num[b] = len(g)
        key = (str(device), anchors.shape[0])
        cache = self.__dict__.setdefault("_gpu_anchor_cache", {})
        if key not in cache:
            cache[key] = torch.from_numpy(anchors).to(device)
        if buffers is None or (buffers.batch, buffers.max_gt, buffers.num_anchors) != (B, max_gt, anchors.shape[0]):
            buffers = ops.AssignBuffers(anchors.shape[0], B, max_gt, device)
        ops.assign_targets(cache[key], torch.from_numpy(gt).to(device), torch.from_numpy(num).to(device), buffers,
                           float(gen.match_threshold), float(gen.unmatch_threshold))
        return dict(labels=buffers.labels, bbox_targets=buffers.bbox_targets, bbox_outside_weights=buffers.bbox_outside_weights,
                    pos_anchor=buffers.pos_anchor, positive_gt_id=buffers.pos_gt_id, num_pos=buffers.num_pos, buffers=buffers)
