"""The reference's CPU path for ONE frame, restated end to end (TEST INFRASTRUCTURE ONLY).

voxelise (oracle.c == point_cloud_ops_v2.py:9-62) -> VFE mean (voxel_encoder.py:205-210) -> sparse encoder (spconv_ref.py; no CPU
implementation of this stage exists in the reference: spconv is GPU / third-party) -> SSFA + head (bev_ref.py == rpn_v1.py:220-235,
mg_head_sessd.py:217-230) -> predict (mg_head_sessd.py:945-1057, rotate_nms_cc restated in oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product.
"""
import time

import numpy as np
import torch

from . import bev_ref, cpu as ocpu, spconv_ref as S

PC_RANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
VOXEL_SIZE = (0.05, 0.05, 0.1)


def layers_to_numpy(layers):
    return [{k: np.asarray(l[k].numpy() if hasattr(l[k], "numpy") else l[k]) for k in ("weight", "gamma", "beta", "mean", "var")}
            for l in layers]


def frame_head(cloud, layers_np, ssfa, head, max_voxels=20000, timings=None):
    """points -> head maps (dict of NHWC tensors) through the CPU oracle."""
    t = timings if timings is not None else {}
    t0 = time.perf_counter()
    v, c, n = ocpu.points_to_voxel(cloud, VOXEL_SIZE, PC_RANGE, 5, max_voxels)
    feat = bev_ref.vfe_mean(torch.from_numpy(v), torch.from_numpy(n)).numpy()
    t["voxelize"] = t.get("voxelize", 0.0) + time.perf_counter() - t0
    t0 = time.perf_counter()
    coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    dense = S.spmiddle_forward(feat, coors, 1, (1408, 1600, 40), layers_np, np.float32)
    t["sparse_encoder"] = t.get("sparse_encoder", 0.0) + time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        neck = bev_ref.ssfa_forward(torch.from_numpy(dense.astype(np.float32)), ssfa)
        hd = bev_ref.head_forward(neck, head)
    t["neck_head"] = t.get("neck_head", 0.0) + time.perf_counter() - t0
    return hd


def frame_detections(cloud, layers_np, ssfa, head, anchors, max_voxels=20000, timings=None, **predict_kw):
    """points -> (boxes [K,7], scores [K], labels [K], aux) with aux['final_anchor'] = anchor index of every returned detection."""
    t = timings if timings is not None else {}
    hd = frame_head(cloud, layers_np, ssfa, head, max_voxels, t)
    t0 = time.perf_counter()
    out = bev_ref.predict_frame(hd["box_preds"].reshape(-1, 7), hd["cls_preds"].reshape(-1), hd["dir_cls_preds"].reshape(-1, 2),
                                hd["iou_preds"].reshape(-1), torch.as_tensor(np.asarray(anchors, np.float32).reshape(-1, 7)),
                                return_aux=True, **predict_kw)
    t["postprocess"] = t.get("postprocess", 0.0) + time.perf_counter() - t0
    return out


def empty_space_logits(ssfa, head):
    """Classification logits (2 anchor types) of the network over empty space: the neck + head on an all-zero BEV map, centre pixel."""
    with torch.no_grad():
        hd = bev_ref.head_forward(bev_ref.ssfa_forward(torch.zeros(1, 128, 48, 48), ssfa), head)
    return hd["cls_preds"][0, 24, 24].numpy().astype(np.float64)
