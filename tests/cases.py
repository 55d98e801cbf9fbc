"""Shared seeded test inputs (the same generators tests/golden/make_golden.py used)."""
import hashlib

import numpy as np

from sessd_data import synth


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def voxel_cases():
    rng = np.random.default_rng(123)
    cases = []
    cases.append(("uniform2k", synth.uniform_cloud(1, 2000), 5, 20000))
    cases.append(("uniform20k", synth.uniform_cloud(0, 20000), 5, 20000))
    cases.append(("ring20k", synth.ring_cloud(0, 20000), 5, 20000))
    cases.append(("cut300", synth.uniform_cloud(2, 5000), 5, 300))
    ctr = rng.uniform([5, -10, -2], [40, 10, 0], (40, 3))
    p = ctr[rng.integers(0, 40, 6000)] + rng.normal(0, 0.04, (6000, 3))
    cases.append(("clustered", np.concatenate([p, rng.uniform(0, 1, (6000, 1))], 1).astype(np.float32), 5, 20000))
    q = synth.uniform_cloud(3, 3000)
    q[::7, 0] = -0.01
    q[1::11, 1] = 40.0
    q[2::13, 2] = 1.0
    q[3::17, 0] = 70.4
    q[4::19] = np.float32([0.0, -40.0, -3.0, 0.5])
    q[5::23, 0] = np.nextafter(np.float32(70.4), np.float32(0))
    cases.append(("edges", q, 5, 20000))
    cases.append(("clustered_mp3_cut", cases[4][1][:4000].copy(), 3, 500))
    cases.append(("empty", np.zeros((0, 4), np.float32), 5, 20000))
    return cases


def iou_inputs():
    b1, _ = synth.random_boxes(11, 160, spread=0.25)
    b2, _ = synth.random_boxes(12, 120, spread=0.25)
    b2[:10] = b1[:10]
    b2[10:15, :2] = b1[10:15, :2] + np.float32([1.6, 0.0]); b2[10:15, 3:7] = b1[10:15, 3:7]
    b1[20, 3] = 0.0
    b1[21:25, 6] = np.float32([0.0, np.pi / 2, -np.pi / 2, np.pi])
    return b1, b2


def assign_cases():
    """(name, gt_boxes [M,7]) inputs of the IoU target assigner (anchors: the 70 400 KITTI car anchors)."""
    out = []
    g12, _ = synth.random_boxes(21, 12)
    g12[:, 2] = -1.0
    out.append(("m12", g12))
    out.append(("m0", np.zeros((0, 7), np.float32)))
    out.append(("m1", synth.random_boxes(22, 1)[0]))
    out.append(("m40", synth.random_boxes(23, 40)[0]))
    e, _ = synth.random_boxes(24, 10)
    e[0, 0] = -20.0                       # no overlap with any anchor => its column max is 0 => never forces a positive
    e[1, 0] = 95.0
    e[2, 3:5] = np.float32([0.5, 0.6])    # pedestrian-sized: max IoU < 0.45, positives only through the forced rule
    e[3, 3:5] = np.float32([0.6, 1.7])
    e[4] = e[5]                           # duplicate GT: argmax ties resolve to the first
    e[6, :2] = np.float32([10.2, 0.2]); e[6, 3:7] = np.float32([1.6, 3.9, 1.56, 0.0])      # exactly an anchor (IoU 1)
    e[7, :2] = np.float32([10.4, 4.4]); e[7, 3:7] = np.float32([1.6, 3.9, 1.56, 0.0])      # half-way between 2 anchors: tie
    e[8, 6] = np.float32(np.pi / 4)       # on the near-bbox swap boundary
    e[9, 6] = np.float32(-3 * np.pi / 4)
    out.append(("edge", e))
    return out


def kitti_wire_case():
    """A KITTI-style calibration (typical values of the training split), image shape and a 5-object annotation dict (1 DontCare)."""
    P2 = np.array([[7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01], [0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01],
                   [0.0, 0.0, 1.0, 2.745884e-03], [0.0, 0.0, 0.0, 1.0]], np.float32)
    R0 = np.eye(4, dtype=np.float32)
    R0[:3, :3] = np.array([[0.9999239, 0.00983776, -0.00744505], [-0.0098698, 0.9999421, -0.00427846],
                           [0.00740253, 0.00435161, 0.9999631]], np.float32)
    Tr = np.eye(4, dtype=np.float32)
    Tr[:3, :4] = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03], [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                           [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01]], np.float32)
    rng = np.random.default_rng(77)
    n = 5
    annos = dict(name=np.array(["Car", "DontCare", "Pedestrian", "Car", "Cyclist"]),
                 location=np.stack([rng.uniform(-10, 10, n), rng.uniform(1.2, 1.9, n), rng.uniform(5, 60, n)], 1),
                 dimensions=np.stack([rng.normal(3.9, 0.3, n), rng.normal(1.56, 0.1, n), rng.normal(1.6, 0.1, n)], 1),
                 rotation_y=rng.uniform(-np.pi, np.pi, n), bbox=rng.uniform(0, 300, (n, 4)), difficulty=np.arange(n, dtype=np.int32))
    info = dict(calib={"P2": P2, "R0_rect": R0, "Tr_velo_to_cam": Tr}, image={"image_shape": np.array([375, 1242], np.int32)}, annos=annos,
                point_cloud={"velodyne_path": "training/velodyne/000007.bin"})
    return info


def head_loss_case():
    """Inputs of the supervised head loss: fused head tensor [2, 35200, 24] (seeded), labels / regression targets of two assigner cases."""
    import torch
    from oracle import anchors as oa
    anc = oa.create_anchors_3d_range().reshape(-1, 7)
    cases = dict(assign_cases())
    labels, targets = [], []
    for name in ("m12", "m40"):
        r = oa.assign_targets(anc, cases[name])
        labels.append(r["labels"])
        targets.append(r["bbox_targets"])
    g = torch.Generator().manual_seed(41)
    head = torch.randn(2, 35200, 24, generator=g) * 0.5
    head[..., 14:16] -= 2.0                       # mostly-negative classification logits, like an early training step
    return head.numpy(), anc, np.stack(labels, 0).astype(np.int32), np.stack(targets, 0).astype(np.float32)


def odiou_pairs():
    """(gboxes, qboxes) [n,7] for the ODIoU loss: predictions = targets + noise at three scales, plus disjoint / contained / identical /
    perpendicular / zero-height-overlap pairs."""
    rng = np.random.default_rng(91)
    g, _ = synth.random_boxes(92, 48)
    q = g.copy()
    for k, s in enumerate((0.02, 0.1, 0.4)):
        sl = slice(16 * k, 16 * (k + 1))
        q[sl] += rng.normal(0, 1, (16, 7)).astype(np.float32) * np.float32([s, s, 0.3 * s, 0.3 * s, 0.5 * s, 0.3 * s, s])
    extra_g = np.float32([[10, 0, -1, 1.6, 3.9, 1.5, 0.3]] * 6)
    extra_q = extra_g.copy()
    extra_q[0, :2] += np.float32([8.0, 6.0])                      # disjoint
    extra_q[1, 3:6] *= np.float32(0.5)                            # contained
    extra_q[2] = extra_g[2]                                       # identical
    extra_q[3, 6] += np.float32(np.pi / 2)                        # perpendicular
    extra_q[4, 2] += np.float32(3.0)                              # no height overlap
    extra_q[5, :2] += np.float32([0.4, -1.1]); extra_q[5, 6] -= np.float32(2.5)
    return np.concatenate([g, extra_g], 0), np.concatenate([q, extra_q], 0)


def checkpoint_model(seed):
    """Small module with the parameter kinds of an SE-SSD checkpoint: a spconv-layout weight [kz,ky,kx,Cin,Cout], BatchNorm1d
    (incl. num_batches_tracked), a Conv2d with bias.  Seeded."""
    import torch
    from torch import nn

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.middle_conv = nn.Module()
            self.middle_conv.weight = nn.Parameter(torch.zeros(3, 3, 3, 4, 16))
            self.bn = nn.BatchNorm1d(16, eps=1e-3, momentum=0.01)
            self.conv_box = nn.Conv2d(8, 14, 1)

    g = torch.Generator().manual_seed(seed)
    m = M()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
        m.bn.running_mean.copy_(torch.randn(16, generator=g))
        m.bn.running_var.copy_(torch.rand(16, generator=g) + 0.5)
        m.bn.num_batches_tracked.fill_(seed)
    return m


def consistency_case():
    """Student / teacher head outputs of two frames for the SE-SSD consistency loss: ~60 objects per frame predicted by both models at
    different anchors (teacher in its own augmentation frame: flip / global rotation / scale undone by the loss), plus unmatched and
    out-of-range predictions.  Returns (preds_stu, preds_tea) as dicts of [2, A, k] float32 arrays, anchors [A, 7], the two transformation
    dicts and the number of planted objects."""
    import torch
    from oracle import anchors as oa
    anc = oa.create_anchors_3d_range().reshape(-1, 7).astype(np.float32)
    A = anc.shape[0]
    rng = np.random.default_rng(2024)
    trans = [dict(flipped=False, noise_rotation=0.031, noise_scale=1.02), dict(flipped=True, noise_rotation=-0.044, noise_scale=0.97)]

    def encode(b, a):                                              # second_box_encode (box_np_ops.py), per pair
        diag = np.sqrt(a[:, 3] ** 2 + a[:, 4] ** 2)
        return np.stack([(b[:, 0] - a[:, 0]) / diag, (b[:, 1] - a[:, 1]) / diag, (b[:, 2] - a[:, 2]) / a[:, 5], np.log(b[:, 3] / a[:, 3]),
                         np.log(b[:, 4] / a[:, 4]), np.log(b[:, 5] / a[:, 5]), b[:, 6] - a[:, 6]], 1).astype(np.float32)

    out = []
    for f in range(2):
        n = 60
        boxes = np.stack([rng.uniform(5, 65, n), rng.uniform(-35, 35, n), rng.uniform(-1.6, -0.6, n), rng.uniform(1.5, 1.8, n),
                          rng.uniform(3.6, 4.4, n), rng.uniform(1.4, 1.7, n), rng.uniform(-3.1, 3.1, n)], 1).astype(np.float32)
        t = trans[f]
        # the same objects in the teacher's frame: undo scale, rotation and flip (inverse of mg_head_sessd.py:668-673)
        tb = boxes.copy()
        tb[:, :6] /= np.float32(t["noise_scale"])
        tb[:, 6] -= np.float32(t["noise_rotation"])
        c, s = np.cos(-t["noise_rotation"]), np.sin(-t["noise_rotation"])
        x, y = tb[:, 0].copy(), tb[:, 1].copy()
        tb[:, 0], tb[:, 1] = x * c + y * s, -x * s + y * c
        if t["flipped"]:
            tb[:, 1] = -tb[:, 1]
            tb[:, 6] = np.float32(np.pi) - tb[:, 6]
        preds = []
        for who, bx in (("stu", boxes), ("tea", tb)):
            g = np.random.default_rng(100 * f + (1 if who == "stu" else 2))
            box = (g.normal(0, 0.05, (A, 7))).astype(np.float32)
            cls = np.full((A, 1), -5.0, np.float32) + g.normal(0, 0.2, (A, 1)).astype(np.float32)
            dr = g.normal(0, 1, (A, 2)).astype(np.float32)
            iou = g.uniform(-1, 1, (A, 1)).astype(np.float32)
            idx = g.choice(A, n + 25, replace=False)
            noisy = bx + g.normal(0, 1, bx.shape).astype(np.float32) * np.float32([0.08, 0.08, 0.03, 0.03, 0.06, 0.03, 0.03])
            noisy[:6] += np.float32([1.5, 1.2, 0, 0, 0, 0, 0.6])       # six objects whose two predictions overlap too little to match
            box[idx[:n]] = encode(noisy, anc[idx[:n]])
            cls[idx[:n], 0] = g.uniform(0.0, 3.0, n).astype(np.float32)
            cls[idx[n:n + 15], 0] = g.uniform(-0.7, 2.0, 15).astype(np.float32)   # confident predictions without a partner
            far = idx[n + 15:]
            box[far, 2] = np.float32(9.0)                                # decoded z above the post-processing range
            cls[far, 0] = np.float32(2.0)
            preds.append(dict(box_preds=box, cls_preds=cls, dir_cls_preds=dr, iou_preds=iou))
        out.append(preds)
    stu = {k: np.stack([out[0][0][k], out[1][0][k]], 0) for k in out[0][0]}
    tea = {k: np.stack([out[0][1][k], out[1][1][k]], 0) for k in out[0][1]}
    return stu, tea, anc, trans
