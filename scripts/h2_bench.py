"""fp16-split tensor-core conv (bevconv_h2.cu): accuracy vs fp64 and timing on the SSFA layer shapes (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import numpy as np, torch
import torch.nn.functional as F
from sessd_b200 import ops
from sessd_b200.runners import _pack_conv


def accuracy():
    g = torch.Generator().manual_seed(0)
    for cin, cout, k, hw in ((128, 128, 3, (21, 37)), (256, 256, 3, (9, 50)), (128, 128, 1, (8, 16)), (128, 24, 1, (20, 33))):
        x = torch.randn(2, cin, hw[0], hw[1], generator=g)
        w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        ref = F.conv2d(x.double(), w.double(), None, 1, k // 2)
        wp, taps = _pack_conv(w)
        taps = [(dy - k // 2, dx - k // 2) for dy, dx in taps]
        xd = x.permute(0, 2, 3, 1).contiguous().cuda()
        out = torch.zeros((2, hw[0], hw[1], cout), device="cuda")
        d = ops.conv_desc(2, hw, cin, hw, cout, hw, taps, relu=False)
        planes, inv = ops.pack_weight_h2(wp.cuda(), 32 if cout <= 32 else -(-cout // 128) * 128)
        amax = torch.zeros(1, device="cuda")
        ops.absmax(xd, amax)
        try:
            ops.bev_conv_h2(xd, planes, inv[:cout].contiguous(), None, None, out, d, amax, None)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("h2 conv failed:", e)
            return False
        got = out.permute(0, 3, 1, 2).cpu().double()
        err = float((got - ref).abs().max() / ref.abs().max())
        print("accuracy cin=%d cout=%d k=%d hw=%s: max rel err %.3e  (amax %.3f)" % (cin, cout, k, hw, err, float(amax[0])))
        if err > 1e-3:
            e = (got - ref).abs()
            idx = np.unravel_index(int(e.argmax()), e.shape)
            print("   worst at", idx, "got", float(got[idx]), "ref", float(ref[idx]))
            print("   sample got/ref:", got[0, :4, 0, 0].numpy(), ref[0, :4, 0, 0].numpy())
    return True


def bench(name, cin, cout, hw_in, hw_out, taps, reps=10, deconv=False):
    x = torch.randn(1, hw_in[0], hw_in[1], cin, device="cuda")
    wp = torch.randn(9 if deconv else len(taps), cin, cout, device="cuda") * 0.05
    cp = -(-cout // 128) * 128 if cout > 32 else 32
    planes, inv = ops.pack_weight_h2(wp, cp)
    wt = ops.pack_weight_tc(wp, cp)
    out = torch.zeros(1, hw_out[0], hw_out[1], cout, device="cuda")
    d = None if deconv else ops.conv_desc(1, hw_in, cin, hw_out, cout, hw_out, taps, relu=True)
    amax = torch.zeros(2, device="cuda")
    ops.absmax(x, amax[0:1])
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    sc = inv[:cout].contiguous()

    def run(kind):
        if deconv:
            if kind == "h2":
                ops.bev_deconv_h2(x, planes, sc, None, None, out, True, amax[0:1], amax[1:2])
            else:
                ops.bev_deconv_tc(x, wt, None, None, None, out)
        elif kind == "h2":
            ops.bev_conv_h2(x, planes, sc, None, None, out, d, amax[0:1], amax[1:2])
        else:
            ops.bev_conv_tc(x, wt, None, None, None, out, d)

    res = {}
    for kind in ("tf32x3", "h2"):
        for _ in range(2):
            run(kind)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(kind); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[kind] = float(np.median(ts)) * 1000
    ntap = 9 if deconv else len(taps)
    fl = 2.0 * hw_in[0] * hw_in[1] * cin * cout * ntap if deconv else 2.0 * hw_out[0] * hw_out[1] * cin * cout * ntap
    print("%-34s GF=%6.2f  us: tf32x3=%6.1f fp16x2=%6.1f   TF/s fp16x2=%6.1f" % (name, fl / 1e9, res["tf32x3"], res["h2"], fl / res["h2"] / 1e6))


if accuracy():
    t3 = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
    bench("conv3x3 128->128 @200x176", 128, 128, (200, 176), (200, 176), t3)
    bench("conv3x3 256->256 @100x88", 256, 256, (100, 88), (100, 88), t3)
    bench("conv1x1 128->128 @200x176", 128, 128, (200, 176), (200, 176), [(0, 0)])
    bench("conv1x1 256->256 @100x88", 256, 256, (100, 88), (100, 88), [(0, 0)])
    bench("deconv 256->128 100x88->200x176", 256, 128, (100, 88), (200, 176), None, deconv=True)
    bench("head 128->24 @200x176", 128, 24, (200, 176), (200, 176), [(0, 0)])

import ctypes
from sessd_b200._lib import lib
print("--- ablations on conv3x3 128->128 @200x176: 1=no split work, 2=no MMAs, 4=no weight reloads, 8=no stores")
x = torch.randn(1, 200, 176, 128, device="cuda"); wp = torch.randn(9, 128, 128, device="cuda") * 0.05
planes, inv = ops.pack_weight_h2(wp, 128); out = torch.zeros(1, 200, 176, 128, device="cuda")
t3 = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
d = ops.conv_desc(1, (200, 176), 128, (200, 176), 128, (200, 176), t3, relu=True)
amax = torch.zeros(1, device="cuda"); ops.absmax(x, amax)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
sc = inv[:128].contiguous()
for mode in (0, 33, 37, 45, 34, 47, 13, 5, 9):
    lib.sessd_set_h2_debug(mode, ctypes.c_void_p(0))
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.bev_conv_h2(x, planes, sc, None, None, out, d, amax, None); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("ablate=%2d  %.1f us" % (mode, float(np.median(ts[2:])) * 1000))
dbg = torch.zeros((1024, 8), dtype=torch.int64, device="cuda")
flush.zero_()
lib.sessd_set_h2_debug(0, ctypes.c_void_p(dbg.data_ptr()))
ops.bev_conv_h2(x, planes, sc, None, None, out, d, amax, None)
torch.cuda.synchronize()
lib.sessd_set_h2_debug(0, ctypes.c_void_p(0))
t = dbg[:148].cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
print("kernel span %.1f us" % ((t[:, 3].max() - t0) / 1000))
dt = (t[:, 3] - t[:, 0]) / 1000
print("   CTA total median %.2f us  (min %.2f max %.2f)" % (np.median(dt), dt.min(), dt.max()))
