"""Seeded model parameters: re-export of sessd_data.weights (the generators live in the library-free package)."""
from sessd_data.weights import *  # noqa: F401,F403
from sessd_data.weights import (SSFA_CONVS, bench_detector_state, kitti_car_anchors, random_detector_state,  # noqa: F401
                                split_detector_state)
