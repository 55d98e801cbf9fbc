"""Model registries, same names as the reference's det3d/models/registry.py so that ``from det3d.models.registry import HEADS`` keeps working."""
from det3d.utils import Registry

_KINDS = dict(READERS="reader", BACKBONES="backbone", NECKS="neck", ROI_EXTRACTORS="roi_extractor", SHARED_HEADS="shared_head",
              HEADS="head", LOSSES="loss", DETECTORS="detector")
globals().update({var: Registry(kind) for var, kind in _KINDS.items()})
__all__ = list(_KINDS)
