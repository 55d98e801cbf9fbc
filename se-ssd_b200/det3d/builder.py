"""String -> object builders used by the config file and the data pipeline (reference: det3d/builder.py:26-35, 38-65, 409-442,
445-500).  ``build_box_coder`` is called INSIDE examples/second/configs/config.py (config.py:5,69)."""
from det3d.core.anchor.anchor_generator import AnchorGeneratorRange
from det3d.core.bbox import region_similarity
from det3d.core.bbox.box_coders import GroundBox3dCoderTorch
from det3d.core.input.voxel_generator import VoxelGenerator


def build_voxel_generator(voxel_config):
    return VoxelGenerator(voxel_size=voxel_config.VOXEL_SIZE, point_cloud_range=voxel_config.RANGE,
                          max_num_points=voxel_config.MAX_POINTS_NUM_PER_VOXEL, max_voxels=20000)


def build_similarity_metric(similarity_config):
    kind = similarity_config.type
    if kind == "nearest_iou_similarity":
        return region_similarity.NearestIouSimilarity()
    raise ValueError("unknown / unsupported similarity type %r (the SE-SSD config uses nearest_iou_similarity)" % kind)


def build_box_coder(box_coder_config):
    kind = box_coder_config["type"]
    if kind == "ground_box3d_coder":
        return GroundBox3dCoderTorch(box_coder_config["linear_dim"], box_coder_config["encode_angle_vector"],
                                     n_dim=box_coder_config.get("n_dim", 9), norm_velo=box_coder_config.get("norm_velo", False))
    raise ValueError("unknown box_coder type")


def build_anchor_generator(anchor_config):
    velocities = anchor_config.velocities if "velocities" in anchor_config else None
    if anchor_config.type == "anchor_generator_range":
        return AnchorGeneratorRange(sizes=anchor_config.sizes, anchor_ranges=anchor_config.anchor_ranges,
                                    rotations=anchor_config.rotations, velocities=velocities,
                                    match_threshold=anchor_config.matched_threshold,
                                    unmatch_threshold=anchor_config.unmatched_threshold, class_name=anchor_config.class_name)
    raise ValueError(" unknown anchor generator type")
