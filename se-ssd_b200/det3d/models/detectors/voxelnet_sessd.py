"""VoxelNet detector (reference: det3d/models/detectors/voxelnet_sessd.py:5-43).  ``forward(example, return_loss=False)`` consumes
the collate_kitti batch dict (voxels, coordinates with batch column, num_points, num_voxels, shape, anchors, calib, metadata) and
returns the per-frame detection dicts, exactly like the reference; every stage runs on the B200 kernels."""
from ..registry import DETECTORS
from .single_stage import SingleStageDetector


@DETECTORS.register_module
class VoxelNet(SingleStageDetector):
    def __init__(self, reader, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__(reader, backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)

    def extract_feat(self, data):
        feats = self.reader(data["voxels"], data["num_points_per_voxel"])
        x = self.backbone(feats, data["coors"], data["batch_size"], data["input_shape"])
        if self.with_neck:
            x = self.neck(x)
        return x

    def forward(self, example, is_ema=[False, None], return_loss=True, **kwargs):
        tag = "_raw" if is_ema[0] else ""
        num_voxels = example["num_voxels" + tag]
        data = dict(voxels=example["voxels" + tag], num_points_per_voxel=example["num_points" + tag],
                    coors=example["coordinates" + tag], batch_size=len(num_voxels), input_shape=example["shape" + tag][0])
        preds = self.bbox_head(self.extract_feat(data))
        if is_ema[0]:
            return preds
        if return_loss:
            return self.bbox_head.loss(example, preds, is_ema[1])
        return self.bbox_head.predict(example, preds, self.test_cfg)
