"""Box layout conversions for the iou3d ops (reference: det3d/core/iou3d/utils.py:74-126)."""

def _wl_index(box_mode, width):
    off = 2 if width == 5 else 3
    return box_mode.index("w") + off, box_mode.index("l") + off


def boxes3d_to_bev_torch(boxes3d, box_mode="wlh", rect=False):
    """[N,7] (x,y,z,dims..,r) or [N,5] -> [N,5] (x1, y1, x2, y2, r): axis-aligned extent before rotation + clockwise angle."""
    if boxes3d.shape[-1] not in (5, 7):
        raise NotImplementedError
    wi, li = _wl_index(box_mode, boxes3d.shape[-1])
    out = boxes3d.new_empty((boxes3d.shape[0], 5))
    hw, hl = boxes3d[:, wi] / 2.0, boxes3d[:, li] / 2.0
    if rect:     # camera coordinates: BEV plane is (x, z)
        cu, cv = boxes3d[:, 0], boxes3d[:, 2]
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = cu - hl, cv - hw, cu + hl, cv + hw
    else:        # velodyne coordinates: BEV plane is (x, y)
        cu, cv = boxes3d[:, 0], boxes3d[:, 1]
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = cu - hw, cv - hl, cu + hw, cv + hl
    out[:, 4] = boxes3d[:, -1]
    return out


def boxes3d_to_bev_3d_torch(boxes3d, box_mode="wlh", rect=False):
    """[N,7] -> [N,7] (x1, y1, z1, x2, y2, z2, r)."""
    wi, li, hi = box_mode.index("w") + 3, box_mode.index("l") + 3, box_mode.index("h") + 3
    out = boxes3d.new_empty((boxes3d.shape[0], 7))
    hw, hl, h = boxes3d[:, wi] / 2.0, boxes3d[:, li] / 2.0, boxes3d[:, hi]
    if rect:
        cu, cv, cw = boxes3d[:, 0], boxes3d[:, 2], boxes3d[:, 1]
        out[:, 0], out[:, 1], out[:, 2] = cu - hl, cv - hw, cw - h
        out[:, 3], out[:, 4], out[:, 5] = cu + hl, cv + hw, cw
    else:
        cu, cv, cw = boxes3d[:, 0], boxes3d[:, 1], boxes3d[:, 2]
        out[:, 0], out[:, 1], out[:, 2] = cu - hw, cv - hl, cw - h / 2.0
        out[:, 3], out[:, 4], out[:, 5] = cu + hw, cv + hl, cw + h / 2.0
    out[:, 6] = boxes3d[:, 6]
    return out
