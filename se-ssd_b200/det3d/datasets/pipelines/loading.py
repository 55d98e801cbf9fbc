"""KITTI wire-format loaders (reference: det3d/datasets/pipelines/loading.py:65-150).  `(res, info) -> (res, info)` transforms that turn a
KITTI info record into the dict the hot path consumes: raw `.bin` points (float32 x, y, z, reflectance), calibration (incl. the image
frustum the detector's post-processing filters with) and GT boxes moved from the camera frame to the velodyne frame."""
from pathlib import Path

import numpy as np

from det3d.core.bbox import box_np_ops

from ..registry import PIPELINES


def remove_dontcare(image_anno):
    """Drop the `DontCare` objects of a KITTI annotation dict (reference: det3d/datasets/kitti/kitti_common.py:506-511)."""
    keep = [i for i, x in enumerate(image_anno["name"]) if x != "DontCare"]
    return {k: v[keep] for k, v in image_anno.items()}


@PIPELINES.register_module
class LoadPointCloudFromFile(object):
    """Points (x, y, z, r in velodyne coordinates) from the (reduced, if present) `.bin` file."""

    def __init__(self, dataset="KittiDataset", **kwargs):
        self.type = dataset
        self.random_select = kwargs.get("random_select", False)
        self.npoints = kwargs.get("npoints", 16834)

    def __call__(self, res, info):
        res["type"] = self.type
        if self.type != "KittiDataset":
            raise NotImplementedError("only the KITTI loader is on the SE-SSD path")
        pc_info = info["point_cloud"]
        velo_path = Path(pc_info["velodyne_path"])
        if not velo_path.is_absolute():
            velo_path = Path(res["metadata"]["image_prefix"]) / pc_info["velodyne_path"]
        reduced = velo_path.parent.parent / (velo_path.parent.stem + "_reduced") / velo_path.name
        if reduced.exists():
            velo_path = reduced
        points = np.fromfile(str(velo_path), dtype=np.float32, count=-1).reshape([-1, res["metadata"]["num_point_features"]])
        res["lidar"]["points"] = points
        return res, info


@PIPELINES.register_module
class LoadPointCloudAnnotations(object):
    """Calibration (+ image frustum) and GT boxes: (x, y, z)cam, l, h, w, ry -> (x, y, z)velo at the box centre, w, l, h, ry."""

    def __init__(self, with_bbox=True, **kwargs):
        self.enable_difficulty_level = kwargs.get("enable_difficulty_level", False)

    def __call__(self, res, info):
        if res["type"] != "KittiDataset":
            raise NotImplementedError("only the KITTI loader is on the SE-SSD path")
        calib = info["calib"]
        res["calib"] = {"rect": calib["R0_rect"], "Trv2c": calib["Tr_velo_to_cam"], "P2": calib["P2"],
                        "frustum": box_np_ops.get_valid_frustum(calib["R0_rect"], calib["Tr_velo_to_cam"], calib["P2"],
                                                                info["image"]["image_shape"])}
        if "annos" in info:
            annos = remove_dontcare(info["annos"])
            gt_boxes = np.concatenate([annos["location"], annos["dimensions"], annos["rotation_y"][..., np.newaxis]], axis=1).astype(np.float32)
            gt_boxes = box_np_ops.box_camera_to_lidar(gt_boxes, calib["R0_rect"], calib["Tr_velo_to_cam"])
            box_np_ops.change_box3d_center_(gt_boxes, [0.5, 0.5, 0], [0.5, 0.5, 0.5])
            res["lidar"]["annotations"] = {"boxes": gt_boxes, "names": annos["name"]}
            if self.enable_difficulty_level:
                res["lidar"]["annotations"]["difficulty"] = annos["difficulty"]
            res.setdefault("cam", {})["annotations"] = {"boxes": annos["bbox"], "names": annos["name"]}
        return res, info
