"""torch-CPU restatement of the supervised SSD-head loss terms of SE-SSD (TEST INFRASTRUCTURE ONLY; gradients through autograd).

Restates, for the terms without the teacher (det3d/models/bbox_heads/mg_head_sessd.py:706-760):
* prepare_loss_weights, NormByNumPositives           mg_head_sessd.py:525-572
* SigmoidFocalLoss(alpha, gamma=2)                   det3d/models/losses/losses.py:345-420
* add_sin_difference + WeightedSmoothL1Loss(sigma)   mg_head_sessd.py:39-44, losses.py:147-204
* get_direction_target + WeightedSoftmaxClassificationLoss   mg_head_sessd.py:62-76, losses.py:489-531
Pinned against the reference's own loss classes on a committed fixture (tests/golden/head_loss_case.npz, make_golden.py `loss`).
"""
import torch


def loss_weights(labels, pos_cls_weight=1.0, neg_cls_weight=1.0, dtype=torch.float32):
    positives, negatives = labels > 0, labels == 0
    cls_w = negatives.type(dtype) * neg_cls_weight + positives.type(dtype) * pos_cls_weight
    reg_w = positives.type(dtype)
    norm = torch.clamp(positives.sum(1, keepdim=True).type(dtype), min=1.0)
    return cls_w / norm, reg_w / norm, labels >= 0


def sigmoid_focal(logits, targets, weights, alpha=0.25, gamma=2.0):
    ce = torch.clamp(logits, min=0) - logits * targets + torch.log1p(torch.exp(-torch.abs(logits)))
    p = torch.sigmoid(logits)
    pt = targets * p + (1 - targets) * (1 - p)
    return torch.pow(1.0 - pt, gamma) * (targets * alpha + (1 - targets) * (1 - alpha)) * ce * weights


def smooth_l1_sin(box_preds, reg_targets, weights, sigma=3.0):
    pe = torch.cat([box_preds[..., :-1], torch.sin(box_preds[..., -1:]) * torch.cos(reg_targets[..., -1:])], -1)
    te = torch.cat([reg_targets[..., :-1], torch.cos(box_preds[..., -1:]) * torch.sin(reg_targets[..., -1:])], -1)
    ad = torch.abs(pe - te)
    lt = (ad <= 1 / sigma ** 2).type_as(ad)
    return (lt * 0.5 * torch.pow(ad * sigma, 2) + (ad - 0.5 / sigma ** 2) * (1.0 - lt)) * weights.unsqueeze(-1)


def direction_ce(dir_logits, anchors, reg_targets, labels, dir_offset=0.0):
    rot_gt = reg_targets[..., -1] + anchors[..., -1]
    tgt = ((rot_gt - dir_offset) > 0).long()
    w = (labels > 0).type_as(dir_logits)
    w = w / torch.clamp(w.sum(-1, keepdim=True), min=1.0)
    ce = torch.nn.functional.cross_entropy(dir_logits.reshape(-1, 2), tgt.reshape(-1), reduction="none").view(w.shape)
    return ce * w


def head_supervised_loss(box_preds, cls_preds, dir_preds, anchors, labels, reg_targets, alpha=0.25, sigma=3.0, dir_offset=0.0,
                         pos_cls_weight=1.0, neg_cls_weight=1.0):
    """box_preds [B,A,7], cls_preds [B,A], dir_preds [B,A,2], anchors [A,7], labels [B,A] int, reg_targets [B,A,7].
    Returns per-frame sums dict(cls, loc, dir, cls_pos, cls_neg) each [B] (the reference divides their batch totals by batch_size)."""
    cls_w, reg_w, cared = loss_weights(labels, pos_cls_weight, neg_cls_weight, box_preds.dtype)
    t = (labels * cared.type_as(labels)).type_as(cls_preds)
    cls = sigmoid_focal(cls_preds, t, cls_w, alpha)
    loc = smooth_l1_sin(box_preds, reg_targets, reg_w, sigma)
    dr = direction_ce(dir_preds, anchors.unsqueeze(0).expand(labels.shape[0], -1, -1), reg_targets, labels, dir_offset)
    return dict(cls=cls.sum(1), loc=loc.sum((1, 2)), dir=dr.sum(1), cls_pos=(cls * (labels > 0)).sum(1), cls_neg=(cls * (labels == 0)).sum(1))


def split_head(head, apl=2):
    """[B, P, stride] fused head tensor (box 2x7 | cls 2 | dir 2x2 | iou 2) -> box [B,A,7], cls [B,A], dir [B,A,2] with A = P * apl."""
    B, P, _ = head.shape
    box = head[..., :7 * apl].reshape(B, P * apl, 7)
    cls = head[..., 7 * apl:8 * apl].reshape(B, P * apl)
    dr = head[..., 8 * apl:10 * apl].reshape(B, P * apl, 2)
    return box, cls, dr


def iou_pred_loss(iou_preds, box_preds, anchors, labels, reg_targets, sigma=3.0):
    """mg_head_sessd.py:755-768: smooth-L1(iou_pred, 2 * aligned_iou3d(decode(pred), decode(target)) - 1) * (1 / num_pos) on the positives,
    per-frame sums [B].  The aligned IoU follows det3d/core/iou3d/iou3d_utils.py:197-252 with the rotated BEV overlap of the C oracle
    (oracle/csrc/oracle.c == the reference's iou3d_cpu.cpp arithmetic).  iou_preds [B,A], box_preds [B,A,7]; the target is a constant."""
    import numpy as np
    from . import bev_ref, cpu as ocpu
    B = labels.shape[0]
    out = []
    for b in range(B):
        pos = labels[b] > 0
        n = int(pos.sum())
        if n == 0:
            out.append(iou_preds[b].sum() * 0.0)
            continue
        q = bev_ref.box_decode(box_preds[b][pos].detach(), anchors[pos]).numpy().astype(np.float32)
        g = bev_ref.box_decode(reg_targets[b][pos], anchors[pos]).numpy().astype(np.float32)
        ov = np.diag(ocpu.boxes_overlap_bev(ocpu.boxes3d_to_bev(q), ocpu.boxes3d_to_bev(g))).astype(np.float32)
        two = np.float32(2)
        lo = np.maximum(q[:, 2] - q[:, 5] / two, g[:, 2] - g[:, 5] / two)
        hi = np.minimum(q[:, 2] + q[:, 5] / two, g[:, 2] + g[:, 5] / two)
        ov3 = ov * np.maximum(hi - lo, np.float32(0))
        iou = ov3 / np.maximum(q[:, 3] * q[:, 4] * q[:, 5] + g[:, 3] * g[:, 4] * g[:, 5] - ov3, np.float32(1e-7))
        target = torch.from_numpy((two * iou - np.float32(1)).astype(np.float32))
        ad = torch.abs(iou_preds[b][pos] - target)
        lt = (ad <= 1 / sigma ** 2).type_as(ad)
        loss = lt * 0.5 * torch.pow(ad * sigma, 2) + (ad - 0.5 / sigma ** 2) * (1.0 - lt)
        out.append((loss / float(max(n, 1))).sum())
    return torch.stack(out)
