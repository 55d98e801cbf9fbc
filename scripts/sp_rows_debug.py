"""Single-layer check of the pair-proportional narrow-layer kernel (spconv_rows.cu) against an fp64 gather-GEMM reference (torch, GPU)
and timing against the dense output-stationary kernels at several neighbour fills; prints, does not assert."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200 import ops

torch.manual_seed(1)
dev = "cuda"


def timeit(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(cin, cout, kvol, n_in, n_out, density, relu=True, reps=0):
    cap_in, cap_out = n_in + 37, n_out + 91
    feat = torch.zeros((cap_in, cin), device=dev)
    feat[:n_in] = torch.relu(torch.randn((n_in, cin), device=dev)) * 3.0
    feat[n_in:] = float("nan")                               # stale rows must never be read
    nbr = torch.full((cap_out, kvol), -1, dtype=torch.int32, device=dev)
    pick = torch.rand((n_out, kvol), device=dev) < density
    idx = torch.randint(0, n_in, (n_out, kvol), device=dev, dtype=torch.int32)
    nbr[:n_out] = torch.where(pick, idx, torch.full_like(idx, -1))
    if n_out > 600:
        nbr[128:256] = -1
        nbr[256:384, 1:] = -1
    w = torch.randn((kvol, cin, cout), device=dev) * 0.1
    sc = torch.rand((cout,), device=dev) + 0.5
    sh = torch.randn((cout,), device=dev) * 0.1
    d_nout = torch.tensor([n_out], dtype=torch.int32, device=dev)
    f64 = torch.cat([feat[:n_in].double(), torch.zeros((1, cin), dtype=torch.float64, device=dev)], 0)
    ref = torch.zeros((n_out, cout), dtype=torch.float64, device=dev)
    nb = nbr[:n_out].long()
    nb = torch.where(nb < 0, torch.full_like(nb, n_in), nb)
    for k in range(kvol):
        ref += f64[nb[:, k]] @ w[k].double()
    ref = ref * sc.double() + sh.double()
    if relu:
        ref = torch.relu(ref)
    out = torch.full((cap_out, cout), -7.0, device=dev)
    amax = torch.zeros((1,), device=dev)
    ops.spconv_forward_rows(feat, nbr, d_nout, cap_out, w, sc, sh, relu, out, amax)
    torch.cuda.synchronize()
    got = out[:n_out].double()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item() / scale
    simt = torch.empty_like(out)
    ops.spconv_forward(feat, nbr, d_nout, cap_out, w, sc, sh, relu, simt)
    same = bool((simt[:n_out] == out[:n_out]).all().item())
    msg = "cin %2d cout %2d kvol %2d n_out %8d fill %.2f : err %.3e  == dense SIMT kernel bitwise %s  untouched %s  amax ok %s" % (
        cin, cout, kvol, n_out, density, err, same, bool((out[n_out:] == -7.0).all().item()),
        amax.item() == out[:n_out].abs().max().item())
    if reps:
        t_rows = timeit(lambda: ops.spconv_forward_rows(feat, nbr, d_nout, cap_out, w, sc, sh, relu, out, amax), reps)
        t_simt = timeit(lambda: ops.spconv_forward(feat, nbr, d_nout, cap_out, w, sc, sh, relu, simt), reps)
        pairs = int((nbr[:n_out] >= 0).sum().item())
        msg += "\n      rows %.4f ms (%.2f TFLOP/s, %.0f GB/s alg) | dense SIMT %.4f ms" % (
            t_rows, 2.0 * pairs * cin * cout / t_rows / 1e9, (4.0 * pairs * cin + 4.0 * n_out * cout + 4.0 * kvol * n_out) / t_rows / 1e6, t_simt)
        if cin >= 16:
            cp = 32
            planes = ops.alloc_planes(cap_in, cp, dev)
            d_nin = torch.tensor([n_in], dtype=torch.int32, device=dev)
            am = torch.zeros((2,), device=dev)
            feat2 = torch.nan_to_num(feat)
            ops.absmax_rows(feat2, d_nin, cap_in, am[0:1]); ops.split_h2(feat2, d_nin, cap_in, am[0:1], planes)
            tiles, inv = ops.pack_weight_sp_h2(w, cp, layout="h2")
            scl = (sc * inv).contiguous()
            t_h2 = timeit(lambda: ops.spconv_forward_h2(planes, am[0:1], nbr, d_nout, cap_out, tiles, scl, sh, relu, simt, am[1:2]), reps)
            msg += " | TMA-gather h2 %.4f ms" % t_h2
    print(msg, flush=True)


for cin, cout in ((4, 16), (16, 16), (16, 32), (32, 32)):
    run(cin, cout, 27, 5000, 4000, 0.3)
run(32, 32, 3, 5000, 4000, 0.7)
run(16, 32, 27, 100, 77, 0.1)
run(32, 32, 27, 300, 129, 1.0)
run(32, 32, 27, 4000, 4000, 0.0)
run(16, 16, 27, 5000, 4000, 0.3, relu=False)
for cin, cout, fill in ((16, 32, 0.04), (32, 32, 0.22), (16, 16, 0.04), (4, 16, 0.04), (32, 32, 0.6)):
    run(cin, cout, 27, 1000000, 1000000, fill, reps=5)
