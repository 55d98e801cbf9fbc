"""Loss registry entries named by the config (config.py:66-78).  The training step (losses, backward, DDP) is a "next" row
(SURVEY.md 8f rank 1); these classes carry the hyper-parameters so that the model builds from the unchanged config and fail
loudly if a training forward is attempted."""
from ..registry import LOSSES


class _HotPathOnlyLoss(object):
    def __init__(self, **kwargs):
        self.cfg = dict(kwargs)

    def __call__(self, *args, **kwargs):
        raise NotImplementedError("%s: the training step is not part of this round's hot path (inference only)" % type(self).__name__)


@LOSSES.register_module
class SigmoidFocalLoss(_HotPathOnlyLoss):
    pass


@LOSSES.register_module
class WeightedSmoothL1Loss(_HotPathOnlyLoss):
    pass


@LOSSES.register_module
class WeightedSoftmaxClassificationLoss(_HotPathOnlyLoss):
    pass
