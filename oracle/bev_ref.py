"""torch-CPU restatement of the dense part of the path (TEST INFRASTRUCTURE ONLY).

* VFE mean                 : det3d/models/readers/voxel_encoder.py:205-210
* SSFA neck                : det3d/models/necks/rpn_v1.py:135-210 (layers), :220-235 (forward); BN eps=1e-3 (:131-132)
* head (4 x 1x1 conv)      : det3d/models/bbox_heads/mg_head_sessd.py:202-215, :217-230
* box decode               : det3d/core/bbox/box_torch_ops.py:81-147
* predict / detections     : det3d/models/bbox_heads/mg_head_sessd.py:893-943, :945-1057
* rotate_nms wrapper       : det3d/core/bbox/box_torch_ops.py:527-548

This is also the "torch-CPU head" timing baseline named by BASELINE.json's north_star.
State-dict key names follow the reference modules so the same tensors can be loaded into the
reference classes (tests/golden/make_golden.py does exactly that to pin this file).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cpu

BN_EPS = 1e-3

# name, kind, cin, cout, k, stride, pad   (rpn_v1.py:135-210; Sequential index of the conv inside its block)
SSFA_CONVS = [
    ("bottom_up_block_0.1", "conv", 128, 128, 3, 1, 1),   # preceded by ZeroPad2d(1) + conv pad 0 == pad 1
    ("bottom_up_block_0.4", "conv", 128, 128, 3, 1, 1),
    ("bottom_up_block_0.7", "conv", 128, 128, 3, 1, 1),
    ("bottom_up_block_1.0", "conv", 128, 256, 3, 2, 1),
    ("bottom_up_block_1.3", "conv", 256, 256, 3, 1, 1),
    ("bottom_up_block_1.6", "conv", 256, 256, 3, 1, 1),
    ("trans_0.0", "conv", 128, 128, 1, 1, 0),
    ("trans_1.0", "conv", 256, 256, 1, 1, 0),
    ("deconv_block_0.0", "deconv", 256, 128, 3, 2, 1),
    ("deconv_block_1.0", "deconv", 256, 128, 3, 2, 1),
    ("conv_0.0", "conv", 128, 128, 3, 1, 1),
    ("w_0.0", "conv", 128, 1, 1, 1, 0),
    ("conv_1.0", "conv", 128, 128, 3, 1, 1),
    ("w_1.0", "conv", 128, 1, 1, 1, 0),
]


def _bn_name(conv_name):
    blk, idx = conv_name.rsplit(".", 1)
    return "%s.%d" % (blk, int(idx) + 1)


def ssfa_random_state(seed, dtype=torch.float32):
    """Seeded SSFA state dict (reference key names)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, kind, cin, cout, k, _s, _p in SSFA_CONVS:
        fan_in = cin * k * k
        shape = (cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)
        sd[name + ".weight"] = (torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)).to(dtype)
        bn = _bn_name(name)
        sd[bn + ".weight"] = (1.0 + 0.1 * torch.randn(cout, generator=g)).to(dtype)
        sd[bn + ".bias"] = (0.1 * torch.randn(cout, generator=g)).to(dtype)
        sd[bn + ".running_mean"] = (0.1 * torch.randn(cout, generator=g)).to(dtype)
        sd[bn + ".running_var"] = (1.0 + 0.2 * torch.rand(cout, generator=g)).to(dtype)
        sd[bn + ".num_batches_tracked"] = torch.tensor(0)
    return sd


def head_random_state(seed, dtype=torch.float32, prefix="tasks.0."):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for nm, cout in (("conv_box", 14), ("conv_cls", 2), ("conv_iou", 2), ("conv_dir", 4)):
        sd[prefix + nm + ".weight"] = (torch.randn((cout, 128, 1, 1), generator=g) * math.sqrt(1.0 / 128)).to(dtype)
        sd[prefix + nm + ".bias"] = (0.1 * torch.randn(cout, generator=g)).to(dtype)
    return sd


def vfe_mean(voxels, num_points):
    """voxel_encoder.py:205-210"""
    return (voxels[:, :, :4].sum(dim=1) / num_points.type_as(voxels).view(-1, 1)).contiguous()


def _cbr(x, sd, name, kind, stride, pad, relu=True):
    w = sd[name + ".weight"].to(x.dtype)
    if kind == "conv":
        y = F.conv2d(x, w, None, stride, pad)
    else:
        y = F.conv_transpose2d(x, w, None, stride, pad, output_padding=1)
    bn = _bn_name(name)
    y = F.batch_norm(y, sd[bn + ".running_mean"].to(x.dtype), sd[bn + ".running_var"].to(x.dtype),
                     sd[bn + ".weight"].to(x.dtype), sd[bn + ".bias"].to(x.dtype), False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def ssfa_forward(x, sd, trace=None):
    """rpn_v1.py:220-235, eval-mode BN."""
    spec = {n: (k, s, p) for n, k, _ci, _co, _ks, s, p in SSFA_CONVS}

    def run(name, t, relu=True):
        k, s, p = spec[name]
        return _cbr(t, sd, name, k, s, p, relu)

    x0 = run("bottom_up_block_0.7", run("bottom_up_block_0.4", run("bottom_up_block_0.1", x)))
    x1 = run("bottom_up_block_1.6", run("bottom_up_block_1.3", run("bottom_up_block_1.0", x0)))
    t0 = run("trans_0.0", x0)
    t1 = run("trans_1.0", x1)
    m0 = run("deconv_block_0.0", t1) + t0
    m1 = run("deconv_block_1.0", t1)
    o0 = run("conv_0.0", m0)
    o1 = run("conv_1.0", m1)
    w0 = run("w_0.0", o0, relu=False)
    w1 = run("w_1.0", o1, relu=False)
    w = torch.softmax(torch.cat([w0, w1], dim=1), dim=1)
    out = o0 * w[:, 0:1] + o1 * w[:, 1:]
    if trace is not None:
        trace.update(x0=x0, x1=x1, t0=t0, t1=t1, m0=m0, m1=m1, o0=o0, o1=o1)
    return out


def head_forward(x, sd, prefix="tasks.0."):
    """mg_head_sessd.py:217-230 -> dict of NHWC tensors."""
    def c(nm):
        return F.conv2d(x, sd[prefix + nm + ".weight"].to(x.dtype), sd[prefix + nm + ".bias"].to(x.dtype)) \
            .permute(0, 2, 3, 1).contiguous()
    return dict(box_preds=c("conv_box"), cls_preds=c("conv_cls"), dir_cls_preds=c("conv_dir"), iou_preds=c("conv_iou"))


def box_decode(enc, anchors):
    """box_torch_ops.py:81-147 (7-dim)."""
    xa, ya, za, wa, la, ha, ra = torch.split(anchors, 1, dim=-1)
    xt, yt, zt, wt, lt, ht, rt = torch.split(enc, 1, dim=-1)
    diagonal = torch.sqrt(la ** 2 + wa ** 2)
    xg = xt * diagonal + xa
    yg = yt * diagonal + ya
    zg = zt * ha + za
    lg = torch.exp(lt) * la
    wg = torch.exp(wt) * wa
    hg = torch.exp(ht) * ha
    rg = rt + ra
    return torch.cat([xg, yg, zg, wg, lg, hg, rg], dim=-1)


def rotate_nms(rbboxes, scores, pre_max_size=1000, post_max_size=100, iou_threshold=0.01):
    """box_torch_ops.py:527-548 with rotate_nms_cc restated in oracle.c."""
    k = min(scores.shape[0], pre_max_size)
    if k == 0:
        return torch.zeros([0], dtype=torch.long)
    # torch.topk: descending; ties broken by lower index (made explicit here)
    order = np.lexsort((np.arange(scores.shape[0]), -scores.numpy().astype(np.float64)))[:k]
    sc = scores.numpy()[order]
    bb = rbboxes.numpy()[order]
    dets = np.concatenate([bb, sc[:, None]], 1).astype(np.float32)
    keep = cpu.rotate_nms_cc(dets, iou_threshold, ge=True)[:post_max_size]
    return torch.from_numpy(order[keep]).long()


def predict_frame(box_enc, cls_logit, dir_logit, iou_pred, anchors, score_thresh=0.3, nms_pre=1000, nms_post=100,
                  nms_thr=0.01, post_range=(0, -40.0, -5.0, 70.4, 40.0, 5.0), frustum_surfaces=None,
                  direction_offset=0.0, return_aux=False):
    """mg_head_sessd.py:945-1057 for one frame.  Inputs: [70400,7], [70400], [70400,2], [70400], [70400,7]."""
    boxes = box_decode(box_enc, anchors)
    dir_labels = torch.max(dir_logit, dim=-1)[1]
    total = torch.sigmoid(cls_logit)
    keep = total >= score_thresh
    scores = total[keep]
    iou = (iou_pred + 1) * 0.5
    scores = scores * torch.pow(iou[keep], 4)
    aux = dict(n_candidates=int(keep.sum()))
    if scores.shape[0] == 0:
        out = (torch.zeros([0, 7]), torch.zeros([0]), torch.zeros([0], dtype=torch.long))
        aux["final_anchor"] = torch.zeros([0], dtype=torch.long)
        return out + (aux,) if return_aux else out
    b = boxes[keep]
    d = dir_labels[keep]
    sel = rotate_nms(b[:, [0, 1, 3, 4, 6]], scores, nms_pre, nms_post, nms_thr)
    aux["nms_selected_anchor"] = torch.nonzero(keep).view(-1)[sel]
    fin = aux["nms_selected_anchor"]
    b, d, s = b[sel], d[sel], scores[sel]
    if frustum_surfaces is not None and b.shape[0] > 0:
        ok = points_in_frustum(b[:, :3].numpy(), frustum_surfaces)
        ok = torch.from_numpy(ok)
        b, d, s, fin = b[ok], d[ok], s[ok], fin[ok]
    if b.shape[0] > 0:
        opp = ((b[:, -1] - direction_offset) > 0) ^ (d.byte() == 1)
        b = b.clone()
        b[:, -1] += torch.where(opp, torch.tensor(np.pi).type_as(b), torch.tensor(0.0).type_as(b))
    pr = torch.tensor(post_range, dtype=b.dtype)
    m = (b[:, :3] >= pr[:3]).all(1) & (b[:, :3] <= pr[3:]).all(1)
    aux["final_anchor"] = fin[m]        # anchor index of every returned detection (bench.py matches detections by it)
    out = (b[m], s[m], torch.zeros(int(m.sum()), dtype=torch.long))
    return out + (aux,) if return_aux else out


def points_in_frustum(points, surfaces):
    """det3d/core/bbox/geometry.py:197-245: inside iff all plane signs < 0.  surfaces [1,6,>=3,3]."""
    surfaces = np.asarray(surfaces)
    sv = surfaces[:, :, :2, :] - surfaces[:, :, 1:3, :]
    normal = np.cross(sv[:, :, 0, :], sv[:, :, 1, :])
    d = -np.einsum("aij,aij->ai", normal, surfaces[:, :, 0, :])
    sign = points @ normal[0].T + d[0][None, :]
    return np.all(sign < 0, axis=1)
