"""Python front-end of the iou3d ops with the reference's names and semantics
(det3d/core/iou3d/iou3d_utils.py:32-52, 143-194, 197-252, 254-306)."""
import torch

import iou3d_cuda
import det3d.core.iou3d.utils as utils


def boxes_iou_bev_gpu(boxes_a, boxes_b, box_mode="wlh", metric="rotate_iou", rect=False):
    """Rotated BEV IoU matrix [M, N] of two [*,7] box lists."""
    if metric != "rotate_iou":
        raise NotImplementedError("only metric='rotate_iou' is provided")
    a = utils.boxes3d_to_bev_torch(boxes_a, box_mode, rect).contiguous()
    b = utils.boxes3d_to_bev_torch(boxes_b, box_mode, rect).contiguous()
    out = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_cuda.boxes_iou_bev_gpu(a, b, out)
    return out


def _height_overlap(boxes_a, boxes_b, h_index, pairwise):
    ha, hb = boxes_a[:, h_index] / 2.0, boxes_b[:, h_index] / 2.0
    shape_a, shape_b = ((-1, 1), (1, -1)) if pairwise else ((-1, 1), (-1, 1))
    lo = torch.max((boxes_a[:, 2] - ha).view(*shape_a), (boxes_b[:, 2] - hb).view(*shape_b))
    hi = torch.min((boxes_a[:, 2] + ha).view(*shape_a), (boxes_b[:, 2] + hb).view(*shape_b))
    return torch.clamp(hi - lo, min=0)


def _iou3d(boxes_a, boxes_b, box_mode, rect, need_bev, pairwise):
    if rect:
        raise NotImplementedError("camera-coordinate boxes are outside the LiDAR hot path")
    wi, li, hi = box_mode.index("w") + 3, box_mode.index("l") + 3, box_mode.index("h") + 3
    a = utils.boxes3d_to_bev_torch(boxes_a, box_mode, rect).contiguous()
    b = utils.boxes3d_to_bev_torch(boxes_b, box_mode, rect).contiguous()
    n, m = boxes_a.shape[0], boxes_b.shape[0]
    if pairwise:
        ov = torch.zeros((n, m), dtype=torch.float32, device=boxes_a.device)
        iou3d_cuda.boxes_overlap_bev_gpu(a, b, ov)
        sb = (1, -1)
    else:
        ov = torch.zeros((n, 1), dtype=torch.float32, device=boxes_a.device)
        iou3d_cuda.boxes_aligned_overlap_bev_gpu(a, b, ov)
        sb = (-1, 1)
    area_a = (boxes_a[:, wi] * boxes_a[:, li]).view(-1, 1)
    area_b = (boxes_b[:, wi] * boxes_b[:, li]).view(*sb)
    iou_bev = ov / torch.clamp(area_a + area_b - ov, min=1e-7)
    ov3d = ov * _height_overlap(boxes_a, boxes_b, hi, pairwise)
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(*sb)
    iou3d = ov3d / torch.clamp(vol_a + vol_b - ov3d, min=1e-7)
    return (iou3d, iou_bev) if need_bev else iou3d


def boxes_iou3d_gpu(boxes_a, boxes_b, box_mode="wlh", rect=False, need_bev=False):
    """3-D IoU matrix [N, M] = BEV overlap x height overlap / union volume."""
    return _iou3d(boxes_a, boxes_b, box_mode, rect, need_bev, pairwise=True)


def boxes_aligned_iou3d_gpu(boxes_a, boxes_b, box_mode="wlh", rect=False, need_bev=False):
    """Row-aligned 3-D IoU [N, 1] of two equally long box lists."""
    assert boxes_a.shape[0] == boxes_b.shape[0]
    return _iou3d(boxes_a, boxes_b, box_mode, rect, need_bev, pairwise=False)


def _nms(boxes_conv, scores, thresh, fn):
    order = scores.sort(0, descending=True)[1]
    boxes_sorted = boxes_conv[order].contiguous()
    keep = torch.LongTensor(boxes_sorted.size(0))
    num = fn(boxes_sorted, keep, thresh)
    return order[keep[:num].to(order.device)].contiguous()


def nms_gpu(boxes, scores, thresh, box_mode="wlh"):
    """Rotated BEV NMS; note the reference converts with rect=True here (iou3d_utils.py:263) -- kept for parity."""
    return _nms(utils.boxes3d_to_bev_torch(boxes, box_mode, rect=True), scores, thresh, iou3d_cuda.nms_gpu)


def nms_3d_gpu(boxes, scores, thresh, box_mode="wlh"):
    return _nms(utils.boxes3d_to_bev_3d_torch(boxes, box_mode, rect=False), scores, thresh, iou3d_cuda.nms_3d_gpu)


def nms_normal_gpu(boxes, scores, thresh):
    return _nms(boxes, scores, thresh, iou3d_cuda.nms_normal_gpu)
