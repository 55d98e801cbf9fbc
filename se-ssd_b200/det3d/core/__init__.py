"""Core operators of the hot path: voxel generation, box math, anchors / targets, rotated IoU / NMS."""
