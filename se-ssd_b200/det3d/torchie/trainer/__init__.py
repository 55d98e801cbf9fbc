from .checkpoint import load_checkpoint, load_state_dict, save_checkpoint, weights_to_cpu

from .trainer_sessd import sigmoid_rampup, update_ema_variables

__all__ = ["load_checkpoint", "load_state_dict", "save_checkpoint", "weights_to_cpu", "sigmoid_rampup", "update_ema_variables"]
