"""Per-role wait cycles of bev_conv_p2_kernel (library built with SESSD_DEFINES=-DSESSD_P2_PROFILE): where the CTA's time goes.
    SESSD_DEFINES=-DSESSD_P2_PROFILE python se-ssd_b200/build.py --force && python scripts/p2_prof.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from sessd_b200 import ops
from sessd_b200._lib import lib
from sessd_b200.runners import _pack_conv
import p2_debug
fn = lib._prod.sessd_set_p2_dbg; fn.restype = None; fn.argtypes = [C.c_void_p]
ops.set_p2_cluster(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
g = torch.Generator().manual_seed(1)
for (cin, cout, hw, k) in ((128, 128, (200, 176), 3), (256, 256, (100, 88), 3), (128, 128, (200, 176), 1)):
    x = torch.randn(1, hw[0], hw[1], cin, generator=g).cuda()
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    wp, taps = _pack_conv(w)
    taps = [(dy - k // 2, dx - k // 2) for dy, dx in taps]
    planes, inv = ops.pack_weight_h2(wp.cuda(), -(-cout // 128) * 128)
    sc = (torch.ones(cout, device="cuda") * inv[:cout]).contiguous()
    d = ops.conv_desc(1, hw, cin, hw, cout, hw, taps, relu=True)
    xp, info = p2_debug.to_planes(x)
    oinfo = torch.zeros(2, device="cuda")
    op = ops.alloc_bev_planes(1, hw[0], hw[1], cout, "cuda")
    dbg = torch.zeros((2 * 148, 8), dtype=torch.int64, device="cuda")
    for it in range(3):
        dbg.zero_()
        fn(C.c_void_p(dbg.data_ptr()))
        ops.bev_conv_p2(xp, info, planes, sc, None, None, None, 30.0, 0.0, None, op, oinfo, d)
        torch.cuda.synchronize()
    fn(C.c_void_p(0))
    m = dbg[:148].double().cpu()
    tl = dbg[148:].double().cpu()
    lead = tl[:, 1] > 0
    tn = ["setup done", "MMA loop start", "MMA loop end", "first acc_full seen", "last epilogue done", "-", "before teardown"]
    print("   timeline (clk since kernel entry, mean over CTAs; MMA rows: leaders only): " + "  ".join(
        "%s %.0f" % (tn[i], (tl[lead, i] if i in (1, 2) else tl[:, i]).mean()) for i in (0, 1, 2, 3, 4, 6)))
    names = ["mma:wait acc_free", "mma:wait patch_full", "mma:wait b_full", "mma:total", "items", "epi:wait acc_full", "wload:wait b_empty", "pload:wait patch_empty"]
    print("conv %dx%d %d->%d k%d" % (hw[0], hw[1], cin, cout, k))
    for i, n in enumerate(names):
        print("   %-26s mean %9.0f  min %9.0f  max %9.0f" % (n, m[:, i].mean(), m[:, i].min(), m[:, i].max()))
