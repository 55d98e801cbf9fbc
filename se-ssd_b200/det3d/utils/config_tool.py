"""Config helpers imported by examples/second/configs/config.py:6 (reference: det3d/utils/config_tool.py:47-57)."""
import numpy as np


def get_downsample_factor(model_config):
    """prod(neck ds strides) / last upsample stride * backbone ds_factor  (8 for the SE-SSD config)."""
    neck = model_config["neck"]
    factor = np.prod(neck.get("ds_layer_strides", [1]))
    ups = neck.get("us_layer_strides", [])
    if len(ups) > 0:
        factor /= ups[-1]
    factor *= model_config["backbone"]["ds_factor"]
    factor = int(factor)
    assert factor > 0
    return factor
