from . import backbones, bbox_heads, detectors, losses, necks, readers  # noqa: F401  (populate the registries)
from .builder import build_backbone, build_detector, build_head, build_loss, build_neck, build_reader
from .registry import BACKBONES, DETECTORS, HEADS, LOSSES, NECKS, READERS

__all__ = ["BACKBONES", "DETECTORS", "HEADS", "LOSSES", "NECKS", "READERS", "build_backbone", "build_detector", "build_head",
           "build_loss", "build_neck", "build_reader"]
