"""numpy restatement of anchor generation and the IoU target assigner (TEST INFRASTRUCTURE ONLY).

BASELINE config #1 (CPU-only, numba/numpy in the reference).  Restates:
* anchors              : det3d/core/bbox/box_np_ops.py:780-833 (create_anchors_3d_range),
                         det3d/core/anchor/anchor_generator.py:108-117
* nearest-BEV IoU      : det3d/core/bbox/region_similarity.py:85-98, box_np_ops.py:354-366 (rbbox2d_to_near_bbox),
                         :619-620 (limit_period), :1007-1046 (iou_jit, eps=0)
* target assignment    : det3d/core/anchor/target_ops_v2.py:11-126 (create_target_np),
                         det3d/core/anchor/target_assigner.py:68-136 (assign_v2, enable_similar_type=True)
* box encoding         : det3d/core/bbox/box_np_ops.py:52-113 (second_box_encode)
Pinned against the reference's own code on committed fixtures (tests/golden/anchors_assign.npz).
"""
import numpy as np


def create_anchors_3d_range(feature_size=(1, 200, 176), anchor_range=(0, -40.0, -1.0, 70.4, 40.0, -1.0),
                            sizes=(1.6, 3.9, 1.56), rotations=(0, 1.57), dtype=np.float32):
    ar = np.array(anchor_range, dtype)
    stride = (ar[3] - ar[0]) / feature_size[2]
    z_centers = np.linspace(ar[2], ar[5], feature_size[0], dtype=dtype)
    y_centers = np.linspace(ar[1], ar[4], feature_size[1], endpoint=False, dtype=dtype) + stride / 2
    x_centers = np.linspace(ar[0], ar[3], feature_size[2], endpoint=False, dtype=dtype) + stride / 2
    rot = np.array(rotations, dtype=dtype)
    sz = np.array(sizes, dtype=dtype).reshape(-1, 3)
    nx, ny, nz, ns, nr = len(x_centers), len(y_centers), len(z_centers), sz.shape[0], len(rot)
    out = np.zeros((nz, ny, nx, ns, nr, 7), dtype)
    out[..., 0] = x_centers[None, None, :, None, None]
    out[..., 1] = y_centers[None, :, None, None, None]
    out[..., 2] = z_centers[:, None, None, None, None]
    out[..., 3:6] = sz[None, None, None, :, None, :]
    out[..., 6] = rot[None, None, None, None, :]
    return out


def limit_period(val, offset=0.5, period=2 * np.pi):
    return val - np.floor(val / period + offset) * period


def rbbox2d_to_near_bbox(rb):
    rots = rb[..., -1]
    r = np.abs(limit_period(rots, 0.5, np.pi))
    cond = (r > np.pi / 4)[..., None]
    c = np.where(cond, rb[:, [0, 1, 3, 2]], rb[:, :4])
    return np.concatenate([c[:, :2] - c[:, 2:] / 2, c[:, :2] + c[:, 2:] / 2], -1)   # center_to_minmax_2d_0_5


def iou_aligned(boxes, query, eps=0.0):
    """iou_jit (box_np_ops.py:1007-1046) as a broadcast with the SAME rounding: under numba the fp32 differences are formed in
    fp32, then `+ eps` (a Python float => float64) promotes the rest of the expression to fp64, and the quotient is rounded
    once into the fp32 output array.  (An all-fp32 evaluation differs from the reference by 1 ulp on ~45 % of the entries.)"""
    b = boxes[:, None, :]
    q = query[None, :, :]
    f8 = np.float64
    box_area = ((q[..., 2] - q[..., 0]).astype(f8) + eps) * ((q[..., 3] - q[..., 1]).astype(f8) + eps)
    iw = (np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0])).astype(f8) + eps
    ih = (np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1])).astype(f8) + eps
    ua = ((b[..., 2] - b[..., 0]).astype(f8) + eps) * ((b[..., 3] - b[..., 1]).astype(f8) + eps) + box_area - iw * ih
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = iw * ih / ua
    return np.where((iw > 0) & (ih > 0), ov, 0).astype(boxes.dtype)


def second_box_encode(boxes, anchors):
    xa, ya, za, wa, la, ha, ra = np.split(anchors, 7, axis=1)
    xg, yg, zg, wg, lg, hg, rg = np.split(boxes, 7, axis=1)
    diagonal = np.sqrt(la ** 2 + wa ** 2)
    return np.concatenate([(xg - xa) / diagonal, (yg - ya) / diagonal, (zg - za) / ha,
                           np.log(wg / wa), np.log(lg / la), np.log(hg / ha), rg - ra], axis=1)


def assign_targets(anchors, gt_boxes, matched=0.6, unmatched=0.45):
    """create_target_np with NearestIouSimilarity, all GT classes collapsed to 1."""
    n = anchors.shape[0]
    labels = np.full((n,), -1, np.int32)
    gt_ids = np.full((n,), -1, np.int32)
    if len(gt_boxes) > 0:
        ov = iou_aligned(rbbox2d_to_near_bbox(anchors[:, [0, 1, 3, 4, 6]]),
                         rbbox2d_to_near_bbox(gt_boxes[:, [0, 1, 3, 4, 6]]))
        a2g_arg = ov.argmax(axis=1)
        a2g_max = ov[np.arange(n), a2g_arg]
        g2a_arg = ov.argmax(axis=0)
        g2a_max = ov[g2a_arg, np.arange(ov.shape[1])]
        g2a_max[g2a_max == 0] = -1
        force = np.where(ov == g2a_max)[0]
        labels[force] = 1
        gt_ids[force] = a2g_arg[force]
        pos = a2g_max >= matched
        labels[pos] = 1
        gt_ids[pos] = a2g_arg[pos]
        bg = np.where(a2g_max < unmatched)[0]
    else:
        bg = np.arange(n)
    fg = np.where(labels > 0)[0]
    if len(gt_boxes) == 0:
        labels[:] = 0
    else:
        labels[bg] = 0
        labels[force] = 1
    targets = np.zeros((n, 7), anchors.dtype)
    if len(gt_boxes) > 0:
        targets[fg] = second_box_encode(gt_boxes[a2g_arg[fg]], anchors[fg])
    weights = np.zeros((n,), anchors.dtype)
    weights[labels > 0] = 1.0
    return dict(labels=labels, bbox_targets=targets, bbox_outside_weights=weights, positive_gt_id=gt_ids[fg])
