"""Torch box ops on the hot path (reference: det3d/core/bbox/box_torch_ops.py:23-147, 527-548).

``rotate_nms`` keeps the reference signature but runs top-k + rotated IoU mask + greedy reduction ON THE GPU
(sessd_rotate_nms) instead of ``dets.cpu().numpy()`` -> boost::geometry on one CPU thread (nms_cpu.h:72-168)."""
import torch

from sessd_b200 import ops


def second_box_encode(boxes, anchors, encode_angle_to_vector=False, smooth_dim=False, norm_velo=False):
    if anchors.shape[-1] != 7 or encode_angle_to_vector or smooth_dim:
        raise NotImplementedError("only the 7-dim log-size encoding of the SE-SSD config is supported")
    xa, ya, za, wa, la, ha, ra = torch.split(anchors, 1, dim=-1)
    xg, yg, zg, wg, lg, hg, rg = torch.split(boxes, 1, dim=-1)
    diag = torch.sqrt(la ** 2 + wa ** 2)
    return torch.cat([(xg - xa) / diag, (yg - ya) / diag, (zg - za) / ha, torch.log(wg / wa), torch.log(lg / la),
                      torch.log(hg / ha), rg - ra], dim=-1)


def second_box_decode(box_encodings, anchors, encode_angle_to_vector=False, bin_loss=False, smooth_dim=False, norm_velo=False):
    if anchors.shape[-1] != 7 or encode_angle_to_vector or smooth_dim:
        raise NotImplementedError("only the 7-dim log-size encoding of the SE-SSD config is supported")
    xa, ya, za, wa, la, ha, ra = torch.split(anchors, 1, dim=-1)
    xt, yt, zt, wt, lt, ht, rt = torch.split(box_encodings, 1, dim=-1)
    diag = torch.sqrt(la ** 2 + wa ** 2)
    return torch.cat([xt * diag + xa, yt * diag + ya, zt * ha + za, torch.exp(wt) * wa, torch.exp(lt) * la,
                      torch.exp(ht) * ha, rt + ra], dim=-1)


def rotate_nms(rbboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """rbboxes [n,5] (x,y,w,l,r), scores [n] (CUDA) -> LongTensor of kept indices (<= post_max_size, best first)."""
    n = int(scores.shape[0])
    if n == 0:
        return torch.zeros([0], dtype=torch.long, device=rbboxes.device)
    pre = n if pre_max_size is None else min(n, int(pre_max_size))
    post = pre if post_max_size is None else int(post_max_size)
    b = rbboxes.detach().float().contiguous()
    s = scores.detach().float().contiguous()
    if not b.is_cuda:
        raise RuntimeError("rotate_nms expects CUDA tensors (there is no CPU fallback)")
    cnt = torch.tensor([n], dtype=torch.int32, device=b.device)
    keep, num = ops.rotate_nms(b, s, cnt, n, pre, min(post, pre), float(iou_threshold), ge=True)
    return keep[: int(num.item())].long()
