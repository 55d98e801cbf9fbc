"""Profile target for ncu: warm up the engine, then run `--frames` eager frames inside a cudaProfilerStart/Stop range.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python scripts/profile_frame.py
    ncu --profile-from-start off --set full --import-source on -k regex:<kernel> -c 3 -o prof python scripts/profile_frame.py
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200 import synth, weights
from sessd_b200.engine import FrameEngine

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--cloud", default="ring")
ap.add_argument("--simt", action="store_true")
a = ap.parse_args()
mk = synth.ring_cloud if a.cloud == "ring" else synth.uniform_cloud
eng = FrameEngine(batch=1, max_points_per_frame=20000, use_tc=not a.simt)
layers, ssfa, head = weights.bench_detector_state(a.cloud, 0)
eng.load_weights(layers, ssfa, head, weights.kitti_car_anchors())
for i in range(3):
    eng.infer([mk(i, 20000)])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(a.frames):
    eng.infer([mk(10 + i, 20000)])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.frames, "frames")
