"""Flatten the per-frame result dict into the keys the collate function and the model consume
(reference: det3d/datasets/pipelines/formating.py:13-86)."""
from ..registry import PIPELINES


@PIPELINES.register_module
class Reformat(object):
    def __init__(self, **kwargs):
        pass

    def __call__(self, res, info):
        lidar = res["lidar"]
        vox = lidar["voxels"]
        bundle = dict(metadata=res["metadata"], points=lidar["points"], voxels=vox["voxels"], shape=vox["shape"],
                      num_points=vox["num_points"], num_voxels=vox["num_voxels"], coordinates=vox["coordinates"],
                      anchors=lidar["targets"]["anchors"])
        if "voxels_raw" in lidar:
            raw = lidar["voxels_raw"]
            bundle.update(points_raw=lidar.get("points_raw"), voxels_raw=raw["voxels"], shape_raw=raw["shape"],
                          num_points_raw=raw["num_points"], num_voxels_raw=raw["num_voxels"], coordinates_raw=raw["coordinates"],
                          anchors_raw=lidar["targets_raw"]["anchors"])
        if res.get("calib") is not None:
            bundle["calib"] = res["calib"]
        if res["mode"] != "test" and "annotations" in lidar:
            bundle["annos"] = lidar["annotations"]
        if res["mode"] == "train" and res.get("labeled", True):
            for k in ("labels", "reg_targets", "reg_weights"):
                if k in lidar["targets"]:
                    bundle[k] = lidar["targets"][k]
        return bundle, info
