"""cfg -> module builders with the reference's names (det3d/models/builder.py): ``build_reader``, ``build_backbone``, ``build_neck``,
``build_roi_extractor``, ``build_shared_head``, ``build_head``, ``build_loss``, ``build_detector``.  A list of configs becomes an
``nn.Sequential`` of the built modules."""
from torch import nn

from det3d.utils import build_from_cfg

from . import registry as _reg


def build(cfg, registry, default_args=None):
    make = lambda c: build_from_cfg(c, registry, default_args)       # noqa: E731
    return nn.Sequential(*map(make, cfg)) if isinstance(cfg, list) else make(cfg)


def _builder_for(registry, doc):
    def _build(cfg):
        return build(cfg, registry)
    _build.__doc__ = doc
    return _build


for _fn, _var in (("build_reader", "READERS"), ("build_backbone", "BACKBONES"), ("build_neck", "NECKS"),
                  ("build_roi_extractor", "ROI_EXTRACTORS"), ("build_shared_head", "SHARED_HEADS"), ("build_head", "HEADS"),
                  ("build_loss", "LOSSES")):
    globals()[_fn] = _builder_for(getattr(_reg, _var), "Build a module registered in %s from its config dict." % _var)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """The detector additionally receives the train / test configs as constructor defaults (tools/test.py builds it this way)."""
    return build(cfg, _reg.DETECTORS, {"train_cfg": train_cfg, "test_cfg": test_cfg})


__all__ = ["build", "build_reader", "build_backbone", "build_neck", "build_roi_extractor", "build_shared_head", "build_head", "build_loss",
           "build_detector"]
