"""Norm-layer factory (reference: det3d/models/utils/norm.py:60-111).  The SE-SSD config sets norm_cfg=None, which the modules
turn into BN1d / BN with eps=1e-3, momentum=0.01; the distributed SyncBN variants of the reference are out of scope."""
from torch import nn

norm_cfg = {"BN": ("bn", nn.BatchNorm2d), "BN1d": ("bn1d", nn.BatchNorm1d), "GN": ("gn", nn.GroupNorm)}


def build_norm_layer(cfg, num_features, postfix=""):
    assert isinstance(cfg, dict) and "type" in cfg
    cfg_ = cfg.copy()
    layer_type = cfg_.pop("type")
    if layer_type not in norm_cfg:
        raise KeyError("Unrecognized norm type {}".format(layer_type))
    abbr, cls = norm_cfg[layer_type]
    assert isinstance(postfix, (int, str))
    requires_grad = cfg_.pop("requires_grad", True)
    cfg_.setdefault("eps", 1e-5)
    if layer_type == "GN":
        assert "num_groups" in cfg_
        layer = cls(num_channels=num_features, **cfg_)
    else:
        layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer
