"""numpy box math used on the hot path and by the anchor target assigner (subset of det3d/core/bbox/box_np_ops.py; the
camera / frustum / legacy helpers of the reference are out of scope).  Host-side, one-off or per-GT work."""
import numpy as np


def second_box_encode(boxes, anchors, encode_angle_to_vector=False, smooth_dim=False, cylindrical=False, norm_velo=False):
    """Residual encoding of (x,y,z,w,l,h,r) boxes against anchors (reference :52-113, 7-dim, log sizes)."""
    if anchors.shape[-1] != 7 or encode_angle_to_vector or smooth_dim:
        raise NotImplementedError("only the 7-dim log-size encoding of the SE-SSD config is supported")
    xa, ya, za, wa, la, ha, ra = np.split(anchors, 7, axis=1)
    xg, yg, zg, wg, lg, hg, rg = np.split(boxes, 7, axis=1)
    diag = np.sqrt(la ** 2 + wa ** 2)
    return np.concatenate([(xg - xa) / diag, (yg - ya) / diag, (zg - za) / ha, np.log(wg / wa), np.log(lg / la),
                           np.log(hg / ha), rg - ra], axis=1)


def second_box_decode(box_encodings, anchors, encode_angle_to_vector=False, smooth_dim=False, norm_velo=False):
    if anchors.shape[-1] != 7 or encode_angle_to_vector or smooth_dim:
        raise NotImplementedError("only the 7-dim log-size encoding of the SE-SSD config is supported")
    xa, ya, za, wa, la, ha, ra = np.split(anchors, 7, axis=-1)
    xt, yt, zt, wt, lt, ht, rt = np.split(box_encodings, 7, axis=-1)
    diag = np.sqrt(la ** 2 + wa ** 2)
    return np.concatenate([xt * diag + xa, yt * diag + ya, zt * ha + za, np.exp(wt) * wa, np.exp(lt) * la, np.exp(ht) * ha,
                           rt + ra], axis=-1)


def limit_period(val, offset=0.5, period=2 * np.pi):
    return val - np.floor(val / period + offset) * period


def center_to_minmax_2d(centers, dims, origin=0.5):
    if origin != 0.5:
        raise NotImplementedError
    return np.concatenate([centers - dims / 2, centers + dims / 2], axis=-1)


def rbbox2d_to_near_bbox(rbboxes):
    """Rotated (x,y,w,l,r) -> nearest axis-aligned box; w/l swapped when |r mod pi| > pi/4 (reference :354-366)."""
    rot = np.abs(limit_period(rbboxes[..., -1], 0.5, np.pi))
    swap = (rot > np.pi / 4)[..., np.newaxis]
    c = np.where(swap, rbboxes[:, [0, 1, 3, 2]], rbboxes[:, :4])
    return center_to_minmax_2d(c[:, :2], c[:, 2:])


def iou_jit(boxes, query_boxes, eps=1.0):
    """Axis-aligned IoU matrix [N,K] (reference :1007-1046), vectorised; zero where the boxes do not overlap.
    Rounding follows the reference under numba: differences in the input dtype, then `+ eps` (a Python float = float64)
    promotes the remaining arithmetic to fp64 and the quotient is rounded once into the output dtype."""
    b = boxes[:, None, :]
    q = query_boxes[None, :, :]
    f8 = np.float64
    area_q = ((q[..., 2] - q[..., 0]).astype(f8) + eps) * ((q[..., 3] - q[..., 1]).astype(f8) + eps)
    iw = (np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0])).astype(f8) + eps
    ih = (np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1])).astype(f8) + eps
    ua = ((b[..., 2] - b[..., 0]).astype(f8) + eps) * ((b[..., 3] - b[..., 1]).astype(f8) + eps) + area_q - iw * ih
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = iw * ih / ua
    return np.where((iw > 0) & (ih > 0), ov, 0).astype(boxes.dtype)


def center_to_corner_box2d(centers, dims, angles=None, origin=0.5):
    """[N,4,2] corners, clockwise from the minimum corner, rotated clockwise for positive angles (reference :512-532)."""
    norm = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=dims.dtype) - np.array(origin, dtype=dims.dtype)
    corners = dims.reshape(-1, 1, 2) * norm.reshape(1, 4, 2)
    if angles is not None:
        s, c = np.sin(angles), np.cos(angles)
        x, y = corners[..., 0].copy(), corners[..., 1].copy()
        corners = np.stack([x * c[:, None] + y * s[:, None], -x * s[:, None] + y * c[:, None]], axis=-1)
    return corners + centers.reshape(-1, 1, 2)


def corner_to_standup_nd(boxes_corner):
    return np.concatenate([boxes_corner.min(axis=1), boxes_corner.max(axis=1)], axis=-1)


def create_anchors_3d_range(feature_size, anchor_range, sizes=(1.6, 3.9, 1.56), rotations=(0, np.pi / 2), velocities=None,
                            dtype=np.float32):
    """Anchor grid [D, H, W, num_sizes, num_rots, 7] (reference :780-833); centres at cell centres of the range."""
    if velocities is not None:
        raise NotImplementedError("velocity anchors are not part of the SE-SSD KITTI config")
    from sessd_b200.weights import kitti_car_anchors
    flat = kitti_car_anchors(tuple(feature_size), tuple(anchor_range), tuple(np.ravel(sizes)), tuple(rotations), dtype)
    ns = int(np.array(sizes).reshape(-1, 3).shape[0])
    return flat.reshape(feature_size[0], feature_size[1], feature_size[2], ns, len(rotations), 7)
