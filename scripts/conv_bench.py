"""Micro-benchmark of the tensor-core conv on the SSFA layer shapes for cluster sizes 1/2/4 (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import numpy as np, torch
from sessd_b200 import ops

def bench(name, cin, cout, hw_in, hw_out, taps, stride=1, reps=10):
    x = torch.randn(1, hw_in[0], hw_in[1], cin, device="cuda")
    wp = torch.randn(len(taps), cin, cout, device="cuda") * 0.05
    wt = ops.pack_weight_tc(wp, -(-cout // 128) * 128 if cout > 32 else 32)
    out = torch.zeros(1, hw_out[0], hw_out[1], cout, device="cuda")
    d = ops.conv_desc(1, hw_in, cin, hw_out, cout, hw_out, taps, in_stride=stride, relu=True)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    res = {}
    for cs in (1, 2, 4):
        ops.set_conv_cluster(cs)
        for _ in range(2):
            ops.bev_conv_tc(x, wt, None, None, None, out, d)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.bev_conv_tc(x, wt, None, None, None, out, d); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[cs] = float(np.median(ts)) * 1000
    fl = 2.0 * hw_out[0] * hw_out[1] * cin * cout * len(taps)
    print("%-34s GF=%6.2f  us: cs1=%6.1f cs2=%6.1f cs4=%6.1f   TF/s(best)=%6.1f" % (name, fl / 1e9, res[1], res[2], res[4], fl / min(res.values()) / 1e6))

t3 = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
bench("conv3x3 128->128 @200x176", 128, 128, (200, 176), (200, 176), t3)
bench("conv3x3 s2 128->256 -> 100x88", 128, 256, (200, 176), (100, 88), t3, stride=2)
bench("conv3x3 256->256 @100x88", 256, 256, (100, 88), (100, 88), t3)
bench("conv1x1 128->128 @200x176", 128, 128, (200, 176), (200, 176), [(0, 0)])
bench("conv1x1 256->256 @100x88", 256, 256, (100, 88), (100, 88), [(0, 0)])
bench("deconv class(4 taps) 256->128", 256, 128, (100, 88), (100, 88), [(0, 0), (0, 1), (1, 0), (1, 1)])
bench("head 128->24 @200x176", 128, 24, (200, 176), (200, 176), [(0, 0)])

print("--- ablations on conv3x3 128->128 @200x176 (cluster 1): 1=no split, 2=hi*hi only, 4=no TMA reloads")
from sessd_b200._lib import lib
ops.set_conv_cluster(1)
x = torch.randn(1, 200, 176, 128, device="cuda"); wp = torch.randn(9, 128, 128, device="cuda") * 0.05
wt = ops.pack_weight_tc(wp, 128); out = torch.zeros(1, 200, 176, 128, device="cuda")
d = ops.conv_desc(1, (200, 176), 128, (200, 176), 128, (200, 176), t3, relu=True)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for mode in (0, 1, 2, 4, 3, 5, 6, 7):
    lib.sessd_set_conv_ablate(mode)
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.bev_conv_tc(x, wt, None, None, None, out, d); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("ablate=%d  %.1f us" % (mode, float(np.median(ts[2:])) * 1000))
lib.sessd_set_conv_ablate(0)
