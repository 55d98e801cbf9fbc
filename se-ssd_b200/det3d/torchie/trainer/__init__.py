from .checkpoint import load_checkpoint, load_state_dict, save_checkpoint, weights_to_cpu

__all__ = ["load_checkpoint", "load_state_dict", "save_checkpoint", "weights_to_cpu"]
