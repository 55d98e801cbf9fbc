"""Step-by-step check of csrc/bevconv_p2.cu against fp64 torch convs, each step in its own subprocess (a trap / hang in one step does
not hide the others), then timings of the 3x3 128->128 @200x176 layer vs the in-kernel-split predecessor.
    python scripts/p2_debug.py            (all steps)        python scripts/p2_debug.py STEP   (one step, in-process)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))

STEPS = ["split", "c1x1_cs1", "c1x1_cs2", "c3x3_cs1", "c3x3_cs2", "c3x3_256_cs2", "stride2_cs1", "stride2_cs2", "deconv_cs1", "deconv_cs2",
         "head_cs1", "head_cs2", "orient_uy", "timing"]


def to_planes(xd):
    import torch
    from sessd_b200 import ops
    info = torch.zeros(2, device="cuda")
    ops.absmax(xd, info[0:1])
    planes = ops.alloc_bev_planes(*xd.shape, "cuda")
    ops.bev_split_planes(xd, info, planes)
    return planes, info


def conv_case(cin, cout, k, stride, hw, cs, b=1, relu=True, resid=False, seed=0):
    import torch, torch.nn.functional as F
    from sessd_b200 import ops
    from sessd_b200.runners import _pack_conv
    g = torch.Generator().manual_seed(seed + cin + cout + k)
    x = torch.randn(b, cin, hw[0], hw[1], generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    sc = 1.0 + 0.1 * torch.randn(cout, generator=g)
    sh = 0.1 * torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), None, stride, k // 2) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    ohw = (ref.shape[2], ref.shape[3])
    res = torch.randn(b, cout, ohw[0], ohw[1], generator=g) if resid else None
    if resid:
        ref = ref + res.double()
    wp, taps = _pack_conv(w)
    taps = [(dy - k // 2, dx - k // 2) for dy, dx in taps]
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.zeros((b, ohw[0], ohw[1], cout), device="cuda")
    d = ops.conv_desc(b, hw, cin, ohw, cout, ohw, taps, in_stride=stride, relu=relu)
    cout_pad = 32 if cout <= 32 else -(-cout // 128) * 128
    planes, inv = ops.pack_weight_h2(wp.cuda(), cout_pad)
    xp, info = to_planes(xd)
    oinfo = torch.zeros(2, device="cuda")
    oplanes = ops.alloc_bev_planes(b, ohw[0], ohw[1], cout, "cuda")
    rd = res.permute(0, 2, 3, 1).contiguous().cuda() if resid else None
    rinfo = None
    if resid:
        rinfo = torch.zeros(2, device="cuda")
        ops.absmax(rd, rinfo[0:1])
    ops.set_p2_cluster(cs)
    ops.bev_conv_p2(xp, info, planes, (sc.cuda() * inv[:cout]).contiguous(), sh.cuda(), rd, rinfo, ops.conv_gain(wp.cuda(), sc.cuda()),
                    float(sh.abs().max()), out, oplanes, oinfo, d)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    back = ops.planes_to_float(oplanes, oinfo).permute(0, 3, 1, 2).cpu().double()
    err_p = float((back - ref).abs().max() / ref.abs().max())
    bad = (got - ref).abs() > 1e-4 * ref.abs().max()
    print("  err fp32-out %.3e  planes-out %.3e  amax ok %s  scale %g  bad elements %d / %d" % (
        err, err_p, float(oinfo[0]) == float(out.abs().max()), float(oinfo[1]), int(bad.sum()), bad.numel()))
    if bad.any():
        idx = bad.nonzero()[:8]
        print("  first bad (b, c, y, x):", idx.tolist())
        ys, xs = bad.any(1)[0].nonzero()[:, 0], bad.any(1)[0].nonzero()[:, 1]
        print("  bad y range %d..%d  x range %d..%d  bad channels %d" % (int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max()), int(bad.any(3).any(2)[0].sum())))
    return err < 5e-6 and err_p < 5e-6


def run_step(step):
    import torch
    from sessd_b200 import ops
    torch.manual_seed(0)
    if step == "split":
        x = torch.randn(2, 9, 11, 64, device="cuda") * 37.0
        p, info = to_planes(x)
        back = ops.planes_to_float(p, info)
        err = float((back - x).abs().max() / x.abs().max())
        print("  split round trip err %.3e scale %g amax %g" % (err, float(info[1]), float(info[0])))
        return err < 3e-7
    if step.startswith("c1x1"):
        return conv_case(128, 128, 1, 1, (16, 8), int(step[-1])) and conv_case(128, 128, 1, 1, (21, 37), int(step[-1]), b=2, resid=True)
    if step.startswith("c3x3_256"):
        return conv_case(256, 256, 3, 1, (9, 50), 2, b=2, resid=True)
    if step.startswith("c3x3"):
        return conv_case(128, 128, 3, 1, (16, 8), int(step[-1])) and conv_case(128, 128, 3, 1, (21, 37), int(step[-1]), b=2, resid=True)
    if step.startswith("stride2"):
        return conv_case(128, 256, 3, 2, (40, 48), int(step[-1])) and conv_case(128, 256, 3, 2, (21, 37), int(step[-1]))
    if step.startswith("head"):
        return conv_case(128, 24, 1, 1, (20, 33), int(step[-1]), relu=False)
    if step == "orient_uy":
        return conv_case(128, 128, 3, 1, (200, 176), 2) and conv_case(256, 256, 3, 1, (100, 88), 2)
    if step.startswith("deconv"):
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(3)
        ok = True
        for hw in ((13, 17), (100, 88)):
            b, cin, cout = 2, 256, 128
            x = torch.randn(b, cin, hw[0], hw[1], generator=g)
            w = torch.randn(cin, cout, 3, 3, generator=g) * (2.0 / (cin * 2.25)) ** 0.5
            sc = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sh = 0.1 * torch.randn(cout, generator=g)
            res = torch.randn(b, cout, 2 * hw[0], 2 * hw[1], generator=g)
            ref = F.relu(F.conv_transpose2d(x.double(), w.double(), None, 2, 1, output_padding=1) * sc.double().view(1, -1, 1, 1)
                         + sh.double().view(1, -1, 1, 1)) + res.double()
            w9 = w.permute(2, 3, 0, 1).reshape(9, cin, cout).contiguous().cuda()
            out = torch.zeros((b, 2 * hw[0], 2 * hw[1], cout), device="cuda")
            xd, rd = x.permute(0, 2, 3, 1).contiguous().cuda(), res.permute(0, 2, 3, 1).contiguous().cuda()
            planes, inv = ops.pack_weight_h2(w9, 128)
            xp, info = to_planes(xd)
            rinfo = torch.zeros(2, device="cuda"); ops.absmax(rd, rinfo[0:1])
            oinfo = torch.zeros(2, device="cuda")
            ops.set_p2_cluster(int(step[-1]))
            ops.bev_deconv_p2(xp, info, planes, (sc.cuda() * inv[:cout]).contiguous(), sh.cuda(), rd, rinfo, ops.conv_gain(w9, sc.cuda()),
                              float(sh.abs().max()), out, None, oinfo, True)
            torch.cuda.synchronize()
            err = float((out.permute(0, 3, 1, 2).cpu().double() - ref).abs().max() / ref.abs().max())
            print("  deconv", hw, "err %.3e" % err)
            ok = ok and err < 5e-6
        return ok
    if step == "timing":
        from sessd_b200.runners import _pack_conv
        g = torch.Generator().manual_seed(1)
        flush = torch.empty((64 * 1024 * 1024,), dtype=torch.float32, device="cuda")
        for (cin, cout, hw, k) in ((128, 128, (200, 176), 3), (256, 256, (100, 88), 3), (128, 128, (200, 176), 1), (128, 24, (200, 176), 1)):
            x = torch.randn(1, hw[0], hw[1], cin, generator=g).cuda()
            w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
            wp, taps = _pack_conv(w)
            taps = [(dy - k // 2, dx - k // 2) for dy, dx in taps]
            cout_pad = 32 if cout <= 32 else -(-cout // 128) * 128
            planes, inv = ops.pack_weight_h2(wp.cuda(), cout_pad)
            sc = (torch.ones(cout, device="cuda") * inv[:cout]).contiguous()
            d = ops.conv_desc(1, hw, cin, hw, cout, hw, taps, relu=True)
            xp, info = to_planes(x)
            oinfo = torch.zeros(2, device="cuda")
            op = ops.alloc_bev_planes(1, hw[0], hw[1], cout, "cuda")
            of = torch.zeros((1, hw[0], hw[1], cout), device="cuda")
            amax = torch.zeros(2, device="cuda"); ops.absmax(x, amax[0:1])
            variants = [("p2 cs1 planes-out", lambda: (ops.set_p2_cluster(1), ops.bev_conv_p2(xp, info, planes, sc, None, None, None, 30.0, 0.0, None, op, oinfo, d))),
                                                ("p2 cs2 planes-out", lambda: (ops.set_p2_cluster(2), ops.bev_conv_p2(xp, info, planes, sc, None, None, None, 30.0, 0.0, None, op, oinfo, d))),
                        ("p2 auto fp32-out", lambda: (ops.set_p2_cluster(0), ops.bev_conv_p2(xp, info, planes, sc, None, None, None, 30.0, 0.0, of, None, oinfo, d)))]
            if cin % 64 == 0 and cout >= 4:
                variants.append(("h2 (in-kernel split)", lambda: ops.bev_conv_h2(x, planes, sc, None, None, of, d, amax[0:1], amax[1:2])))
            for name, fn in variants:
                ts = []
                for i in range(13):
                    flush.zero_()
                    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b_.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b_))
                t = sorted(ts[3:])[len(ts[3:]) // 2]
                fl = 2.0 * hw[0] * hw[1] * cin * cout * k * k
                print("  %dx%d %d->%d k%d  %-22s %.1f us  %.0f TFLOP/s algorithmic" % (hw[0], hw[1], cin, cout, k, name, t * 1000, fl / t / 1e9))
        return True
    raise SystemExit("unknown step " + step)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        ok = run_step(sys.argv[1])
        print("  ->", "OK" if ok else "FAIL")
        sys.exit(0 if ok else 1)
    res = {}
    for st in STEPS:
        print("== step", st, flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), st], timeout=int(os.environ.get("P2_STEP_TIMEOUT", "120")), capture_output=True, text=True)
            print(r.stdout[-3000:], end="")
            if r.returncode != 0:
                print(r.stderr[-1500:])
            res[st] = "OK" if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
        except subprocess.TimeoutExpired:
            res[st] = "TIMEOUT"
        print("   [%s, %.1fs]" % (res[st], time.time() - t0), flush=True)
    print("SUMMARY", res)
