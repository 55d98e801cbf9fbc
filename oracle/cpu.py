"""ctypes front-end of oracle/csrc/oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
        _lib.oracle_box_overlap.restype = C.c_float
        _lib.oracle_iou_bev.restype = C.c_float
        _lib.oracle_iou_3d.restype = C.c_float
    return _lib


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def grid_size(voxel_size, pc_range):
    """det3d/core/input/voxel_generator.py:12-16 (fp32 arithmetic, np.round)."""
    r = np.asarray(pc_range, np.float32)
    v = np.asarray(voxel_size, np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def points_to_voxel(points, voxel_size, pc_range, max_points=5, max_voxels=20000):
    """point_cloud_ops_v2.py:120-194 -> (voxels [M,max_points,F], coors [M,3] zyx, num_points [M])."""
    points = np.ascontiguousarray(points, np.float32)
    n, f = points.shape
    vs = np.asarray(voxel_size, np.float32)
    rg = np.asarray(pc_range, np.float32)
    grid = grid_size(voxel_size, pc_range).astype(np.int32)
    voxels = np.zeros((max_voxels, max_points, f), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    m = lib().oracle_points_to_voxel(_p(points), n, f, _p(vs), _p(rg), _p(grid, C.c_int), max_points, max_voxels,
                                     _p(voxels), _p(coors, C.c_int), _p(num, C.c_int))
    assert m >= 0
    return voxels[:m], coors[:m], num[:m]


def _mat(fn, a, b, width):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, width)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, width)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    fn(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def boxes_overlap_bev(a, b):
    return _mat(lib().oracle_boxes_overlap_bev, a, b, 5)


def boxes_iou_bev(a, b):
    return _mat(lib().oracle_boxes_iou_bev, a, b, 5)


def boxes_iou_3d(a, b):
    return _mat(lib().oracle_boxes_iou_3d, a, b, 7)


def nms_sorted(boxes, thresh, mode=0):
    """iou3d nms (boxes pre-sorted by descending score). mode 0 rot-bev, 1 3d, 2 axis-aligned."""
    w = 7 if mode == 1 else 5
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, w)
    keep = np.zeros((boxes.shape[0] + 1,), np.int64)
    k = lib().oracle_nms_sorted(_p(boxes), boxes.shape[0], C.c_float(thresh), mode, _p(keep, C.c_int64))
    return keep[:k]


def rotate_nms_cc(dets, thresh, ge=True):
    """nms_cpu.py:37-48 semantics; dets [n,6] = x,y,w,l,r,score -> keep indices (into dets)."""
    dets = np.ascontiguousarray(dets, np.float32).reshape(-1, 6)
    keep = np.zeros((dets.shape[0] + 1,), np.int64)
    k = lib().oracle_rotate_nms_cc(_p(dets), dets.shape[0], C.c_float(thresh), 1 if ge else 0, _p(keep, C.c_int64))
    return keep[:k]


def corners_standup(dets5):
    dets5 = np.ascontiguousarray(dets5, np.float32).reshape(-1, 5)
    n = dets5.shape[0]
    corners = np.zeros((n, 4, 2), np.float32)
    su = np.zeros((n, 4), np.float32)
    for i in range(n):
        lib().oracle_corners_standup(_p(dets5[i]), _p(corners[i]), _p(su[i]))
    return corners, su


def box_decode(enc, anchors):
    enc = np.ascontiguousarray(enc, np.float32).reshape(-1, 7)
    anchors = np.ascontiguousarray(anchors, np.float32).reshape(-1, 7)
    out = np.zeros_like(enc)
    lib().oracle_box_decode(_p(enc), _p(anchors), enc.shape[0], _p(out))
    return out


def boxes3d_to_bev(boxes7):
    """det3d/core/iou3d/utils.py:74-101, velo coords, box_mode 'wlh'."""
    b = np.asarray(boxes7, np.float32)
    out = np.zeros((b.shape[0], 5), np.float32)
    hw, hl = b[:, 3] / np.float32(2), b[:, 4] / np.float32(2)
    out[:, 0], out[:, 1] = b[:, 0] - hw, b[:, 1] - hl
    out[:, 2], out[:, 3] = b[:, 0] + hw, b[:, 1] + hl
    out[:, 4] = b[:, 6]
    return out


def boxes3d_to_bev3d(boxes7):
    """det3d/core/iou3d/utils.py:104-126 (velo coords)."""
    b = np.asarray(boxes7, np.float32)
    out = np.zeros((b.shape[0], 7), np.float32)
    hw, hl, h = b[:, 3] / np.float32(2), b[:, 4] / np.float32(2), b[:, 5]
    out[:, 0], out[:, 1], out[:, 2] = b[:, 0] - hw, b[:, 1] - hl, b[:, 2] - h / np.float32(2)
    out[:, 3], out[:, 4], out[:, 5] = b[:, 0] + hw, b[:, 1] + hl, b[:, 2] + h / np.float32(2)
    out[:, 6] = b[:, 6]
    return out
