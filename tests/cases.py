"""Shared seeded test inputs (the same generators tests/golden/make_golden.py used)."""
import hashlib

import numpy as np

from sessd_b200 import synth


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
