"""numpy-level rotated NMS entry point (reference: det3d/ops/nms/nms_cpu.py:37-48).  Same signature and return type (a python
list of kept indices into ``dets``); the O(n^2) polygon work runs on the GPU (sessd_rotate_nms), not in boost::geometry."""
import numpy as np
import torch

from sessd_b200 import ops


def rotate_nms_cc(dets, thresh):
    """dets [n,6] = (x, y, w, l, r, score) -> kept indices, best first."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return []
    d = torch.from_numpy(dets).cuda()
    cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
    keep, num = ops.rotate_nms(d[:, :5].contiguous(), d[:, 5].contiguous(), cnt, n, n, n, float(thresh), ge=True)
    return keep[: int(num.item())].cpu().tolist()
