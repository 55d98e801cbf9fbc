"""IoU target assignment (reference: det3d/core/anchor/target_ops_v2.py:11-126).  Host numpy, per frame; BASELINE config #1.
A GPU version is a "next" row."""
import numpy as np


def create_target_np(all_anchors, gt_boxes, similarity_fn, box_encoding_fn, prune_anchor_fn=None, gt_classes=None,
                     matched_threshold=0.6, unmatched_threshold=0.45, bbox_inside_weight=None, positive_fraction=None,
                     rpn_batch_size=300, norm_by_num_examples=False, box_code_size=7):
    n = all_anchors.shape[0]
    if gt_classes is None:
        gt_classes = np.ones([gt_boxes.shape[0]], dtype=np.int32)
    labels = np.full((n,), -1, dtype=np.int32)
    gt_ids = np.full((n,), -1, dtype=np.int32)
    have_gt = len(gt_boxes) > 0
    if have_gt:
        overlap = similarity_fn(all_anchors, gt_boxes)                 # [n, m]
        best_gt = overlap.argmax(axis=1)
        best_gt_iou = overlap[np.arange(n), best_gt]
        best_anchor = overlap.argmax(axis=0)
        best_anchor_iou = overlap[best_anchor, np.arange(overlap.shape[1])]
        best_anchor_iou[best_anchor_iou == 0] = -1                     # GTs matching nothing are dropped
        # every anchor tying a GT's best IoU is a forced positive
        forced = np.where(overlap == best_anchor_iou)[0]
        forced_gt = best_gt[forced]
        labels[forced] = gt_classes[forced_gt]
        gt_ids[forced] = forced_gt
        pos = best_gt_iou >= matched_threshold
        labels[pos] = gt_classes[best_gt[pos]]
        gt_ids[pos] = best_gt[pos]
        bg = np.where(best_gt_iou < unmatched_threshold)[0]
    else:
        bg = np.arange(n)
    fg = np.where(labels > 0)[0]
    fg_iou = best_gt_iou[fg] if have_gt else None
    if have_gt:
        labels[bg] = 0
        labels[forced] = gt_classes[forced_gt]                          # forced positives survive the background pass
    else:
        labels[:] = 0
    targets = np.zeros((n, box_code_size), dtype=all_anchors.dtype)
    if have_gt:
        targets[fg, :] = box_encoding_fn(gt_boxes[best_gt[fg], :], all_anchors[fg, :])
    weights = np.zeros((n,), dtype=all_anchors.dtype)
    if norm_by_num_examples:
        weights[labels > 0] = 1.0 / np.maximum(1.0, np.sum(labels >= 0))
    else:
        weights[labels > 0] = 1.0
    return {"labels": labels, "bbox_targets": targets, "bbox_outside_weights": weights, "assigned_anchors_overlap": fg_iou,
            "positive_gt_id": gt_ids[fg], "assigned_anchors_inds": fg}
