from . import box_coders, box_np_ops, box_torch_ops, geometry, region_similarity

__all__ = ["box_coders", "box_np_ops", "box_torch_ops", "geometry", "region_similarity"]
