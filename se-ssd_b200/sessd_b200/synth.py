"""Seeded synthetic inputs: re-export of sessd_data.synth (the generators live in the library-free package)."""
from sessd_data.synth import *  # noqa: F401,F403
from sessd_data.synth import PC_RANGE, VOXEL_SIZE, random_boxes, ring_cloud, uniform_cloud  # noqa: F401
