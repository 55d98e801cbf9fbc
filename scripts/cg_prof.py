"""Per-role cycle counters of spconv_cg_kernel per layer at the stress / frame shape (library built with SESSD_DEFINES=-DSESSD_CG_PROFILE).
    SESSD_DEFINES=-DSESSD_CG_PROFILE python se-ssd_b200/build.py --force && python scripts/cg_prof.py [stress|frame] [deep]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import numpy as np, torch
from sessd_data import synth, weights
from sessd_b200 import ops
from sessd_b200._lib import lib
from sessd_b200.engine import FrameEngine
shape = sys.argv[1] if len(sys.argv) > 1 else "stress"
lib.sessd_set_sp_cg_deep(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
fn = lib._prod.sessd_set_cg_dbg; fn.restype = None; fn.argtypes = [C.c_void_p]
if shape == "stress":
    B, N = 16, 200000
    clouds = [synth.uniform_cloud(1000 + f, N) for f in range(B)]
    eng = FrameEngine(batch=B, max_points_per_frame=N, max_voxels=200000, growth=(1.0, 8.0, 8.0, 8.0, 8.0))
    kind = "uniform"
else:
    B, N = 1, 20000
    clouds = [synth.ring_cloud(0, N)]
    eng = FrameEngine(batch=1, max_points_per_frame=N)
    kind = "ring"
layers, ssfa, head = weights.bench_detector_state(kind, 0)
eng.load_weights(layers, ssfa, head, weights.kitti_car_anchors())
eng.stage(clouds)
dbg = torch.zeros((296, 16), dtype=torch.int64, device="cuda")
names = ["mma wait full_b", "mma wait full_a", "mma loop", "tiles", "wload wait empty", "prod wait empty", "prod loop", "epi wait acc", "epi total", "list build", "kernel", "fills"]
rows = []
def mark(label):
    if label.startswith("conv:"):
        torch.cuda.synchronize()
        rows.append((label, dbg.double().cpu().clone()))
        dbg.zero_()
with torch.cuda.stream(eng.stream):
    eng.d_points.copy_(eng.h_points, non_blocking=True); eng.d_off.copy_(eng.h_off, non_blocking=True)
    ops.voxelize(eng.d_points, eng.d_off, eng.vox)
    for it in range(2):
        rows.clear(); dbg.zero_()
        fn(C.c_void_p(dbg.data_ptr()))
        eng.sparse_and_neck(mark=mark)
        torch.cuda.synchronize()
fn(C.c_void_p(0))
for label, m in rows:
    li = int(label.split(":")[1])
    if eng.middle.plan[li]["impl"] != "cg":
        continue
    act = m[:, 3] > 0
    if not act.any():
        continue
    mm = m[act]
    t = mm[:, 3].mean()
    print("%s  cin %d cout %d  CTAs %d  tiles/CTA %.1f  fills/tile %.1f  kernel %.0f clk" % (label, eng.middle.plan[li]["cin"], eng.middle.plan[li]["cout"], int(act.sum()), t, mm[:, 11].mean() / t, mm[:, 10].mean()))
    print("    per tile (clk): " + "  ".join("%s %.0f" % (names[i], mm[:, i].mean() / t) for i in (9, 2, 0, 1, 6, 5, 4, 8, 7)))
