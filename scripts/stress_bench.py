"""BASELINE config #5 shape on ONE GPU: dense-scene stress -- 200k-point uniform clouds, 0.05 m voxels, batch 16 frames per launch,
max_voxels 200000.  Reports per-stage device times and the achieved HBM GB/s of the bandwidth-bound kernels against their
ALGORITHMIC bytes: voxeliser 16N + 112M, dense scatter 4*B*H*W*C; checks frame 0 against the CPU oracle."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import numpy as np, torch
from sessd_b200 import ops, synth, weights
from sessd_b200.engine import FrameEngine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--points", type=int, default=200000)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
B, N = a.batch, a.points
clouds = [synth.uniform_cloud(1000 + f, N) for f in range(B)]
eng = FrameEngine(batch=B, max_points_per_frame=N, max_voxels=200000, growth=(1.0, 8.0, 8.0, 8.0, 8.0))
layers, ssfa, head = weights.split_detector_state(weights.random_detector_state(0, cls_bias=-3.0))
eng.load_weights(layers, ssfa, head, weights.kitti_car_anchors())
eng.calibrate_cls_bias(clouds, 400)
torch.cuda.synchronize()
print("GPU memory allocated: %.1f GB" % (torch.cuda.memory_allocated() / 2 ** 30))
eng.stage(clouds)
st = eng.stream
res = []
for it in range(a.iters + 1):
    with torch.cuda.stream(st):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        eng.d_points.copy_(eng.h_points, non_blocking=True); eng.d_off.copy_(eng.h_off, non_blocking=True)
        ev[0].record(st)
        ops.voxelize(eng.d_points, eng.d_off, eng.vox)
        ev[1].record(st)
        dense = eng.middle.forward(eng.vox.mean, eng.vox.coors, eng.vox.num_voxels[B:B + 1])
        ev[2].record(st)
        _, hd = eng.neck.forward(dense)
        ev[3].record(st)
        ops.postprocess(hd, eng.anchors, None, eng.post)
        ev[4].record(st)
        st.synchronize()
    if it:
        res.append([ev[k].elapsed_time(ev[k + 1]) for k in range(4)])
r = np.median(np.array(res), 0)
nv = eng.vox.num_voxels.cpu().numpy()
M = int(nv[B])
sites = [int(l["n"].item()) for l in eng.middle.levels[1:]]
status = int(eng.middle.status.item())
vox_bytes = 16.0 * N * B + 112.0 * M
line = {"config": "stress: uniform-%dk x batch %d, 1 GPU" % (N // 1000, B), "frames_per_sec": B / (r.sum() / 1000), "stage_ms": dict(zip(
    ("voxelize", "sparse_encoder", "neck_head", "postprocess"), [round(float(x), 3) for x in r])), "voxels": M, "active_sites": sites,
    "capacity_status": status, "voxelize_alg_bytes": vox_bytes, "voxelize_GBps": vox_bytes / (r[0] / 1000) / 1e9,
    "detections_frame0": int(eng.post.count[0].item())}
# parity spot check of frame 0 against the oracle
from oracle import cpu as ocpu
ov, oc, on = ocpu.points_to_voxel(clouds[0], synth.VOXEL_SIZE, synth.PC_RANGE, 5, 200000)
m0 = int(nv[0])
line["frame0_voxel_parity"] = bool(m0 == len(oc) and np.array_equal(eng.vox.coors[:m0, 1:].cpu().numpy(), oc)
                                   and np.array_equal(eng.vox.num_points[:m0].cpu().numpy(), on))
print(json.dumps(line))
