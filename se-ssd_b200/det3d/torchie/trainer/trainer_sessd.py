"""Teacher / student bookkeeping of the SE-SSD trainer (reference: det3d/torchie/trainer/trainer_sessd.py:250-275, 302-318).  Only the
pieces that sit on the per-step path are mirrored: the consistency-weight ramp and the teacher's exponential moving average (one kernel
launch over the flat parameter arenas, sessd_b200.train)."""
import numpy as np

from sessd_b200 import train as _train


def sigmoid_rampup(current_epoch, max_epochs=1, rampup_length=15.0):
    """consistency-weight ramp exp(-5 (1 - t)^2), t = clip(epoch, 0, 15) / 15 (trainer_sessd.py:302-309); 1.0 when max_epochs == 0"""
    if max_epochs == 0:
        return 1.0
    phase = 1.0 - float(np.clip(current_epoch, 0.0, rampup_length)) / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def update_ema_variables(model, ema_model, global_step):
    """ema = alpha ema + (1 - alpha) param with alpha = min(1 - 1 / (global_step + 1), 0.999) (trainer_sessd.py:315-318); ``model`` and
    ``ema_model`` are ParamArenas (or modules carrying one as ``._arena``)."""
    a = getattr(model, "_arena", model)
    b = getattr(ema_model, "_arena", ema_model)
    return _train.update_ema_variables(a, b, global_step)
