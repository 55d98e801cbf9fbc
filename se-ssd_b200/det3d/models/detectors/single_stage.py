"""Single-stage detector glue (reference: det3d/models/detectors/single_stage.py:9-39)."""
from torch import nn

from .. import builder
from ..registry import DETECTORS


@DETECTORS.register_module
class SingleStageDetector(nn.Module):
    def __init__(self, reader, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.reader = builder.build_reader(reader)
        self.backbone = builder.build_backbone(backbone)
        if neck is not None:
            self.neck = builder.build_neck(neck)
        self.bbox_head = builder.build_head(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    def forward_dummy(self, example):
        return self.bbox_head(self.extract_feat(example))
