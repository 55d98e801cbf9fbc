"""``collate_kitti``: list of per-frame dicts -> batch dict = the wire format of the GPU stage
(reference: det3d/torchie/parallel/collate.py:154-218): voxels / num_points / num_voxels concatenated, ``coordinates`` gets a
leading batch-index column, anchors stacked per task, calib stacked, metadata kept as a list."""
import collections

import numpy as np
import torch

_CAT = {"voxels", "num_points", "num_gt", "voxel_labels", "num_voxels"}
_COOR = {"coordinates", "points"}
_PER_TASK = {"anchors", "anchors_mask", "reg_targets", "reg_weights", "labels"}


def _base(key):
    return key[:-4] if key.endswith("_raw") else key


def collate_kitti(batch_list, samples_per_gpu=1):
    merged = collections.defaultdict(list)
    for example in batch_list:
        for k, v in example.items():
            merged[k].append(v)
    out = {}
    for key, elems in merged.items():
        base = _base(key)
        if base in _CAT:
            out[key] = torch.tensor(np.concatenate(elems, axis=0))
        elif base in _COOR:
            out[key] = torch.tensor(np.concatenate(
                [np.pad(c, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, c in enumerate(elems)], axis=0))
        elif base in _PER_TASK:
            per_task = collections.defaultdict(list)
            for elem in elems:
                for t, arr in enumerate(elem):
                    per_task[t].append(torch.tensor(arr))
            out[key] = [torch.stack(per_task[t]) for t in sorted(per_task)]
        elif key == "metadata":
            out[key] = elems
        elif key == "calib":
            out[key] = {}
            for elem in elems:
                for k1, v1 in elem.items():
                    out[key].setdefault(k1, []).append(v1)
            out[key] = {k1: torch.tensor(np.stack(v1, axis=0)) for k1, v1 in out[key].items()}
        elif key == "gt_boxes":
            ntask = len(elems[0])
            res = []
            for t in range(ntask):
                m = max(len(e[t]) for e in elems)
                buf = np.zeros((len(elems), m, 7))
                for i, e in enumerate(elems):
                    buf[i, : len(e[t])] = e[t]
                res.append(buf)
            out[key] = res
        else:
            out[key] = np.stack(elems, axis=0)
    return out
