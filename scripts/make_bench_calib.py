#!/usr/bin/env python
"""Generate se-ssd_b200/sessd_data/bench_calib.json: the classification calibration of the bench / parity workload.

Random-init weights give a BEV map whose empty regions are constant, so with a plain bias shift hundreds of anchors share
bit-identical logits around the score threshold and the kept set depends on tie-breaking (the round-1 bench parity leg measured
mkldnn thread blocking, not the CUDA path).  Here, per cloud kind, the CPU oracle computes on the seed-0 frame
    empty_logit[a] : the network's classification logit over empty space, per anchor type
    alpha          : logit' = alpha (logit - empty_logit[a]) + EMPTY_LOGIT puts ~400 anchors over the 0.3 threshold
and the constants are COMMITTED, so both bench arms and the tests load bit-identical weights (no per-arm calibration).
CPU only (oracle); run from the repo root:  python scripts/make_bench_calib.py
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "se-ssd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import frame as oframe  # noqa: E402
from sessd_data import synth, weights  # noqa: E402

TARGET = 400


def neck_gains(dense, ssfa):
    """Per-layer BatchNorm gain that brings the rms of the signal (pre-activation minus its silent-input value) over the responding
    pixels to 1, layer by layer in forward order, on the calibration frame.  Rounded to 3 significant digits: the committed constants
    are the definition."""
    from oracle import bev_ref
    spec = {n: (k, st, p) for n, k, _ci, _co, _ks, st, p in bev_ref.SSFA_CONVS}
    gains = {}
    acts = {}

    def run(name, x, sd):
        k, st, p = spec[name]
        return bev_ref._cbr(x, sd, name, k, st, p, relu=False)

    def calibrated(name, x):
        sd = weights.quiet_neck_state(ssfa, 0, gains)
        y = run(name, x, sd)
        shift = sd[weights._bn_of(name) + ".bias"].view(1, -1, 1, 1)
        sig = y - shift
        act = sig != 0
        rms = float(torch.sqrt((sig[act] ** 2).mean()))
        gains[name] = float("%.3g" % (1.0 / rms))
        sd = weights.quiet_neck_state(ssfa, 0, gains)
        return torch.relu(run(name, x, sd))

    with torch.no_grad():
        x = torch.from_numpy(dense.astype(np.float32))
        a = calibrated("bottom_up_block_0.1", x)
        a = calibrated("bottom_up_block_0.4", a)
        x0 = calibrated("bottom_up_block_0.7", a)
        a = calibrated("bottom_up_block_1.0", x0)
        a = calibrated("bottom_up_block_1.3", a)
        x1 = calibrated("bottom_up_block_1.6", a)
        t0 = calibrated("trans_0.0", x0)
        t1 = calibrated("trans_1.0", x1)
        m0 = calibrated("deconv_block_0.0", t1) + t0
        m1 = calibrated("deconv_block_1.0", t1)
        acts["o0"] = calibrated("conv_0.0", m0)
        acts["o1"] = calibrated("conv_1.0", m1)
    return gains


def main():
    from oracle import bev_ref, cpu as ocpu, spconv_ref as S
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    out = {"seed": 0}
    dense = {}
    layers, ssfa0, _ = weights.split_detector_state(weights.random_detector_state(0))
    lnp = oframe.layers_to_numpy(layers)
    for kind in ("ring", "uniform"):
        cloud = synth.ring_cloud(0, 20000) if kind == "ring" else synth.uniform_cloud(0, 20000)
        v, c, n = ocpu.points_to_voxel(cloud, oframe.VOXEL_SIZE, oframe.PC_RANGE, 5, 20000)
        feat = bev_ref.vfe_mean(torch.from_numpy(v), torch.from_numpy(n)).numpy()
        dense[kind] = S.spmiddle_forward(feat, np.concatenate([np.zeros((len(c), 1), np.int32), c], 1), 1, (1408, 1600, 40), lnp, np.float32)
    out["neck_gains"] = neck_gains(dense["ring"], ssfa0)
    print("neck gains", out["neck_gains"])
    for kind in ("ring", "uniform"):
        x = torch.from_numpy(dense[kind].astype(np.float32))

        def logits(alpha):
            _, ssfa, head = weights.bench_detector_state(kind, 0, alpha=alpha, gains=out["neck_gains"])
            with torch.no_grad():
                return bev_ref.head_forward(bev_ref.ssfa_forward(x, ssfa), head)["cls_preds"].reshape(-1).numpy(), ssfa, head

        lg, ssfa, head = logits(1.0)
        e = oframe.empty_space_logits(ssfa, head)
        assert np.all(e == weights.EMPTY_LOGIT), e                             # the quiet neck is exactly silent over empty space
        d = np.sort(lg.astype(np.float64) - weights.EMPTY_LOGIT)[::-1]
        kth = 0.5 * (d[TARGET - 1] + d[TARGET])
        assert kth > 0, "fewer than %d anchors respond" % TARGET
        alpha = float("%.4g" % ((math.log(0.3 / 0.7) - weights.EMPTY_LOGIT) / kth))     # rounded: the constant is the definition
        lg, _, _ = logits(alpha)
        n_cand = int((torch.sigmoid(torch.from_numpy(lg)).numpy() >= 0.3).sum())
        top = np.sort(lg)[::-1][:1000]
        out[kind] = dict(alpha=alpha, candidates_on_seed0=n_cand, distinct_in_top1000=int(len(np.unique(top))),
                         top_logit=float(top[0]), n_above_empty=int((lg > weights.EMPTY_LOGIT).sum()),
                         n_at_empty=int((lg == weights.EMPTY_LOGIT).sum()))
        print(kind, out[kind])
    path = os.path.join(ROOT, "se-ssd_b200", "sessd_data", "bench_calib.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
