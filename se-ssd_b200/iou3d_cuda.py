"""``iou3d_cuda`` -- drop-in for the reference's torch extension of the same (top-level) name
(det3d/core/iou3d/src/iou3d.cpp:268-281, imported by det3d/core/iou3d/iou3d_utils.py:2).  Same ten entry points, same
argument conventions: inputs are contiguous CUDA float32 tensors (else an error is raised, mirroring CHECK_INPUT :7-9),
outputs are allocated by the caller, ``nms_*`` fill a CPU int64 ``keep`` tensor and return the number kept, everything else
returns 1.  Work is done by libsessd_b200.so (csrc/iou3d.cu) on torch's CURRENT stream -- the reference launches on the legacy
default stream with a blocking cudaMalloc/cudaMemcpy inside nms_gpu (:131-142)."""
import torch

from sessd_b200 import ops


def _chk(t, name, cuda=True):
    if cuda and not t.is_cuda:
        raise RuntimeError("%s must be a CUDAtensor " % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous " % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _chk(t, n)
    ops.boxes_overlap_bev(boxes_a, boxes_b, ans_overlap)
    return 1


def boxes_aligned_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _chk(t, n)
    ops.boxes_aligned_overlap_bev(boxes_a, boxes_b, ans_overlap)
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _chk(t, n)
    ops.boxes_iou_bev(boxes_a, boxes_b, ans_iou)
    return 1


def boxes_iou3d_gpu(boxes_a, boxes_b, ans_iou):
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _chk(t, n)
    ops.boxes_iou3d(boxes_a, boxes_b, ans_iou)
    return 1


def _nms(boxes, keep, thresh, mode):
    _chk(boxes, "boxes")
    if not keep.is_contiguous() or keep.dtype != torch.int64:
        raise RuntimeError("keep must be a contiguous int64 tensor")
    k, num = ops.nms_sorted(boxes, float(thresh), mode)
    n = int(num.item())
    keep[:n] = k[:n].to(keep.device)
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 0)


def nms_3d_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 1)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 2)


# The reference's *_cpu entry points (iou3d_cpu.cpp) exist for host tensors.  This build has no CPU path: they stage the
# inputs through the GPU kernels and copy the result back (same values up to libm-vs-CUDA sin/cos/atan2 ulps).
def _via_gpu(fn, boxes_a, boxes_b, out):
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (out, "out")):
        _chk(t, n, cuda=False)
    tmp = torch.zeros(out.shape, dtype=torch.float32, device="cuda")
    fn(boxes_a.cuda(), boxes_b.cuda(), tmp)
    out.copy_(tmp)
    return 1


def boxes_overlap_bev_cpu(boxes_a, boxes_b, ans_overlap):
    return _via_gpu(ops.boxes_overlap_bev, boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    return _via_gpu(ops.boxes_iou_bev, boxes_a, boxes_b, ans_iou)


def boxes_iou3d_cpu(boxes_a, boxes_b, ans_iou):
    return _via_gpu(ops.boxes_iou3d, boxes_a, boxes_b, ans_iou)
