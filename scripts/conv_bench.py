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
    for cs in (1, 2, 4, 8):
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
    print("%-34s GF=%6.2f  us: cs1=%6.1f cs2=%6.1f cs4=%6.1f cs8=%6.1f  TF/s(best)=%6.1f" % (name, fl / 1e9, res[1], res[2], res[4], res[8], fl / min(res.values()) / 1e6))

t3 = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
from sessd_b200._lib import lib
import sys as _s
variant = int(_s.argv[1]) if len(_s.argv) > 1 else 1
lib.sessd_set_conv_variant(variant)
print("=== conv variant", variant)
bench("conv3x3 128->128 @200x176", 128, 128, (200, 176), (200, 176), t3)
bench("conv3x3 s2 128->256 -> 100x88", 128, 256, (200, 176), (100, 88), t3, stride=2)
bench("conv3x3 256->256 @100x88", 256, 256, (100, 88), (100, 88), t3)
bench("conv1x1 128->128 @200x176", 128, 128, (200, 176), (200, 176), [(0, 0)])
bench("conv1x1 256->256 @100x88", 256, 256, (100, 88), (100, 88), [(0, 0)])
bench("deconv class(4 taps) 256->128", 256, 128, (100, 88), (100, 88), [(0, 0), (0, 1), (1, 0), (1, 1)])
bench("head 128->24 @200x176", 128, 24, (200, 176), (200, 176), [(0, 0)])

print("--- ablations on conv3x3 128->128 @200x176 (cluster 1): 1=no split, 2=hi*hi only, 4=no TMA reloads")
ops.set_conv_cluster(1)
lib.sessd_set_conv_variant(variant)
x = torch.randn(1, 200, 176, 128, device="cuda"); wp = torch.randn(9, 128, 128, device="cuda") * 0.05
wt = ops.pack_weight_tc(wp, 128); out = torch.zeros(1, 200, 176, 128, device="cuda")
d = ops.conv_desc(1, (200, 176), 128, (200, 176), 128, (200, 176), t3, relu=True)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for mode in ((0, 4) if variant == 2 else (0, 1, 2, 4, 7)):
    lib.sessd_set_conv_ablate(mode)
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.bev_conv_tc(x, wt, None, None, None, out, d); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("ablate=%d  %.1f us" % (mode, float(np.median(ts[2:])) * 1000))
lib.sessd_set_conv_ablate(0)

print("--- per-CTA timeline (globaltimer ns): setup, first TMA, mainloop, epilogue")
import ctypes
for name, cin, cout, hw, taps in (("conv3x3 128->128 @200x176", 128, 128, (200, 176), t3), ("conv1x1 128->128 @200x176", 128, 128, (200, 176), [(0, 0)])):
    x = torch.randn(1, hw[0], hw[1], cin, device="cuda"); wp = torch.randn(len(taps), cin, cout, device="cuda") * 0.05
    wt = ops.pack_weight_tc(wp, 128); out = torch.zeros(1, hw[0], hw[1], cout, device="cuda")
    d = ops.conv_desc(1, hw, cin, hw, cout, hw, taps, relu=True)
    dbg = torch.zeros((1024, 8), dtype=torch.int64, device="cuda")
    for i in range(2):
        ops.bev_conv_tc(x, wt, None, None, None, out, d)
    torch.cuda.synchronize()
    lib.sessd_set_conv_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    ops.bev_conv_tc(x, wt, None, None, None, out, d)
    torch.cuda.synchronize()
    lib.sessd_set_conv_debug_buffer(ctypes.c_void_p(0))
    t = dbg[:275].cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    print(name, "kernel span %.1f us" % ((t[:, 4].max() - t0) / 1000))
    for lbl, a, b in (("setup", 0, 1), ("first TMA", 1, 2), ("mainloop", 2, 3), ("epilogue", 3, 4), ("CTA total", 0, 4)):
        dt = (t[:, b] - t[:, a]) / 1000
        print("   %-10s median %.2f us  (min %.2f max %.2f)" % (lbl, np.median(dt), dt.min(), dt.max()))
    st = (t[:, 0] - t0) / 1000
    print("   CTA start times: first wave <1us: %d, later: median %.1f us" % (int((st < 1).sum()), float(np.median(st[st >= 1])) if (st >= 1).any() else 0))

    if variant == 1:
        tr = dbg.view(-1)[4096:4096 + 36 * 8].cpu().numpy().reshape(36, 8).astype(np.float64)
        base = tr[0, 0]
        print("   step trace CTA0 (us since first producer issue): P=producer got empty, Cf=conv got full, Cd=conv done, Mf=mma got full, Ms=mma got split, Mi=mma issued")
        for it in list(range(0, min(12, len(taps) * (cin // 32)))):
            r = (tr[it, :6] - base) / 1000
            print("   it=%2d  P=%6.2f Cf=%6.2f Cd=%6.2f Mf=%6.2f Ms=%6.2f Mi=%6.2f" % (it, r[0], r[1], r[2], r[3], r[4], r[5]))
