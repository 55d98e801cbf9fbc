from .norm import build_norm_layer

__all__ = ["build_norm_layer"]
