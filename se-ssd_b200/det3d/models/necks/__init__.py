from .rpn_v1 import SSFA

__all__ = ["SSFA"]
