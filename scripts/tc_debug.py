"""Diagnostics for the tcgen05 conv kernel: identity-weight 1x1 convs with index-coded inputs reveal layout mistakes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import numpy as np, torch
from sessd_b200 import ops

from sessd_b200._lib import lib as _L

def suite():
    def run(x_nhwc, wp, cout, taps, tc=True):
        b, h, w, cin = x_nhwc.shape
        out = torch.full((b, h, w, cout), -777.0, device="cuda")
        d = ops.conv_desc(b, (h, w), cin, (h, w), cout, (h, w), taps, relu=False)
        if tc:
            ops.bev_conv_tc(x_nhwc.cuda(), ops.pack_weight_tc(wp.cuda(), 32 if cout <= 32 else 128), None, None, None, out, d)
        else:
            ops.bev_conv(x_nhwc.cuda(), wp.cuda(), None, None, None, out, d)
        torch.cuda.synchronize()
        return out.cpu()

    h, w, c = 8, 16, 128
    eye = torch.eye(c).reshape(1, c, c)
    # A1: pixel-coded input
    x = torch.zeros(1, h, w, c)
    pix = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
    x[0] = pix[..., None].expand(h, w, c)
    o = run(x, eye, c, [(0, 0)])
    print("A1 pixel-coded: exact =", bool(torch.equal(o, x)), " max|err| =", float((o - x).abs().max()))
    if not torch.equal(o, x):
        print(" out[0,:2,:8,0] =", o[0, :2, :8, 0].tolist()); print(" out[0,0,0,:40] =", o[0, 0, 0, :40].tolist())
    # A2: channel-coded input
    x = torch.arange(c, dtype=torch.float32).reshape(1, 1, 1, c).expand(1, h, w, c).contiguous()
    o = run(x, eye, c, [(0, 0)])
    print("A2 channel-coded: exact =", bool(torch.equal(o, x)), " max|err| =", float((o - x).abs().max()))
    if not torch.equal(o, x):
        print(" out[0,0,0,:40] =", o[0, 0, 0, :40].tolist()); print(" out[0,3,5,:40] =", o[0, 3, 5, :40].tolist())
    # A3: fine mantissa (needs the lo terms): values with 20 significant bits
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, h, w, c, generator=g)
    o = run(x, eye, c, [(0, 0)])
    print("A3 random x, identity W: max rel err =", float((o - x).abs().max() / x.abs().max()))
    wr = torch.randn(1, c, c, generator=g) / c ** 0.5
    ref = (x.double().reshape(-1, c) @ wr[0].double()).reshape(1, h, w, c)
    o = run(x, wr, c, [(0, 0)])
    print("A4 random x, random W: max rel err =", float((o.double() - ref).abs().max() / ref.abs().max()))
    o2 = run(x, wr, c, [(0, 0)], tc=False)
    print("A4 (SIMT) max rel err =", float((o2.double() - ref).abs().max() / ref.abs().max()))
    # A5: tap shift (dy,dx) = (1,-1) with identity weights => shifted copy with zero fill
    xs = torch.randn(1, 16, 32, c, generator=g)
    o = run(xs, eye, c, [(1, -1)])
    ref = torch.zeros_like(xs); ref[:, :-1, 1:] = xs[:, 1:, :-1]
    print("A5 tap shift: max err =", float((o - ref).abs().max()))

    # A6: cluster multicast of the weight tiles (2 / 4 CTAs per cluster) and the strided-TMA stride-2 conv
    import torch.nn.functional as F
    xs = torch.randn(2, 37, 45, c, generator=g)
    w3 = torch.randn(128, c, 3, 3, generator=g) * (2.0 / (c * 9)) ** 0.5
    ref = F.conv2d(xs.permute(0, 3, 1, 2).double(), w3.double(), None, 1, 1).permute(0, 2, 3, 1)
    wp3 = w3.permute(2, 3, 1, 0).reshape(9, c, 128).contiguous()
    taps = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
    for cs in (1, 2, 4):
        ops.set_conv_cluster(cs)
        try:
            o = run(xs, wp3, 128, taps)
            print("A6 cluster=%d conv3x3: max rel err =" % cs, float((o.double() - ref).abs().max() / ref.abs().max()))
        except Exception as e:
            print("A6 cluster=%d FAILED:" % cs, repr(e)[:200]); break
    ops.set_conv_cluster(1)
    ref2 = F.conv2d(xs.permute(0, 3, 1, 2).double(), w3.double(), None, 2, 1).permute(0, 2, 3, 1)
    b_, h2, w2 = ref2.shape[0], ref2.shape[1], ref2.shape[2]
    out = torch.full((b_, h2, w2, 128), -777.0, device="cuda")
    d = ops.conv_desc(b_, (37, 45), c, (h2, w2), 128, (h2, w2), taps, in_stride=2, relu=False)
    try:
        ops.bev_conv_tc(xs.cuda(), ops.pack_weight_tc(wp3.cuda(), 128), None, None, None, out, d)
        torch.cuda.synchronize()
        print("A7 stride-2 conv (strided TMA): max rel err =", float((out.cpu().double() - ref2).abs().max() / ref2.abs().max()))
    except Exception as e:
        print("A7 stride-2 FAILED:", repr(e)[:200])


for variant in (1, 2, 3):
    print('========== conv variant', variant)
    _L.sessd_set_conv_variant(variant)
    try:
        suite()
    except Exception as e:
        print('variant', variant, 'FAILED:', repr(e)[:300])
        break
_L.sessd_set_conv_variant(1)

# E1: which rounding does the tensor core apply to raw fp32 bits fed as tf32?  (variant 1, A_hi left raw)
print("========== E1: raw fp32 as tf32 operand")
import torch.nn.functional as F
g = torch.Generator().manual_seed(3)
xs = torch.rand(1, 32, 32, 128, generator=g) + 0.5          # positive data: truncation bias would show up clearly
w3 = torch.randn(128, 128, 3, 3, generator=g) * (2.0 / (128 * 9)) ** 0.5
ref = F.conv2d(xs.permute(0, 3, 1, 2).double(), w3.double(), None, 1, 1).permute(0, 2, 3, 1)
wp3 = w3.permute(2, 3, 1, 0).reshape(9, 128, 128).contiguous()
taps = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
for mode, name in ((0, "hi written back (reference)"), (8, "raw hi, lo = a - trunc(a)"), (16, "raw hi, lo = a - rna(a)")):
    _L.sessd_set_conv_ablate(mode)
    out = torch.zeros(1, 32, 32, 128, device="cuda")
    d = ops.conv_desc(1, (32, 32), 128, (32, 32), 128, (32, 32), taps, relu=False)
    ops.bev_conv_tc(xs.cuda(), ops.pack_weight_tc(wp3.cuda(), 128), None, None, None, out, d)
    torch.cuda.synchronize()
    print("E1 %-32s max rel err = %.3e" % (name, float((out.cpu().double() - ref).abs().max() / ref.abs().max())))
_L.sessd_set_conv_ablate(0)
