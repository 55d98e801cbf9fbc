import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200._lib import lib
out = torch.zeros(16, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    rc = lib.sessd_latency_probe(2000, ctypes.c_void_p(out.data_ptr()), st)
    torch.cuda.synchronize()
o = out.cpu().tolist()
names = ["commit(empty)->own wait", "mbarrier 2-warp round trip", "tcgen05.st x32 + wait::st", "1 MMA(N256 f16 TS)+commit->wait",
         "4 MMAs+commit->wait", "tcgen05.ld x32 + wait::ld", "commit->other warp->arrive back", "8 commits back-to-back -> wait", "tcgen05.st x32+wait under MMA load", "tcgen05.ld x32+wait under MMA load", "pair N256->[0,256) + N128->[128,256) (overlap)", "pair N256 + N128->[256,384) (disjoint)", "single N256 MMA"]
print("rc", rc)
for n, v in zip(names, o):
    print("%-36s %6d clk" % (n, v))
