"""Loss registry entries named by the config (config.py:66-78).  They carry the reference's hyper-parameters (same attribute names:
``_alpha``, ``_gamma``, ``_sigma``, ``_loss_weight``, ...) so that the model builds from the unchanged config.  The fused device
implementation of the three supervised terms is ``MultiGroupHead.loss_supervised`` (csrc/headloss.cu); calling a loss object
element-wise like the reference's python modules is not supported (the remaining training step is a "next" row)."""
from ..registry import LOSSES


class _FusedOnDevice(object):
    def __call__(self, *args, **kwargs):
        raise NotImplementedError("%s: use MultiGroupHead.loss_supervised (fused value + gradient on the device)" % type(self).__name__)


@LOSSES.register_module
class SigmoidFocalLoss(_FusedOnDevice):
    """reference: det3d/models/losses/losses.py:365-420"""

    def __init__(self, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        if float(gamma) != 2.0:
            raise NotImplementedError("the fused focal loss is built for gamma = 2 (config.py:72)")
        self._alpha, self._gamma, self._reduction, self._loss_weight = alpha, gamma, reduction, loss_weight


@LOSSES.register_module
class WeightedSmoothL1Loss(_FusedOnDevice):
    """reference: det3d/models/losses/losses.py:147-204 (code_weights are ignored there as well: `_code_weights = None`)"""

    def __init__(self, sigma=3.0, reduction="mean", code_weights=None, codewise=True, loss_weight=1.0):
        self._sigma, self._code_weights, self._codewise, self._reduction, self._loss_weight = sigma, None, codewise, reduction, loss_weight


@LOSSES.register_module
class WeightedSoftmaxClassificationLoss(_FusedOnDevice):
    """reference: det3d/models/losses/losses.py:498-531"""

    def __init__(self, logit_scale=1.0, loss_weight=1.0, name=""):
        if float(logit_scale) != 1.0:
            raise NotImplementedError("logit_scale != 1 is not used by the SE-SSD config")
        self.name, self._loss_weight, self._logit_scale = name, loss_weight, logit_scale
