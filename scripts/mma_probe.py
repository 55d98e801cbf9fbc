import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200._lib import lib
out = torch.zeros(8, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for n, modes in ((128, (0, 4, 8, 16, 20)), (256, (0, 1, 4, 5))):
    for mode in modes:
        for _ in range(2):
            rc = lib.sessd_mma_probe(n, 4096, mode, ctypes.c_void_p(out.data_ptr()), st)
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        cyc = o[1] / 4096
        print("N=%3d mode=%2d [bit2: commit/12 MMAs, bit3: 2 commits/12, bit4: try_wait+fence/12] (A from %s, %s acc): %.1f clk/MMA issue, %.1f clk/MMA retire -> %.0f MAC/clk/SM, %.0f TFLOP/s chip @%.2f GHz" % (
            n, mode, "TMEM" if mode & 1 else "smem", "2 rotating" if mode & 2 else "1", o[0] / 4096, cyc, 128 * n * 8 / cyc,
            2 * 128 * n * 8 / cyc * 148 * (o[1] / max(o[2], 1)) / 1000, o[1] / max(o[2], 1)))
