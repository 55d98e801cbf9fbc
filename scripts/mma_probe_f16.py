"""Cycles per tcgen05.mma kind::f16 (M128 x N x K16, SS) vs N, accumulator rotation and swizzle mode (lab library)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200._lib import lib, check
out = torch.zeros(4, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 2000
for sw in (0, 4):
    for n in (32, 64, 128, 256):
        for rot in (0, 1, 2):
            if rot == 2 and n > 128:
                continue
            for rep in range(2):
                check(lib.sessd_mma_probe_f16(n, iters, rot | sw, C.c_void_p(out.data_ptr()), st), "probe")
                torch.cuda.synchronize()
            o = out.cpu().tolist()
            print("swizzle %-4s N %3d accumulators %d: issue %.1f clk/MMA, retire %.1f clk/MMA (floor %d)" % ("64B" if sw else "128B", n, 1 << rot, o[0] / iters, o[1] / iters, 128 * n // 256))
