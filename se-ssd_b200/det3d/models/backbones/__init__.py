"""Sparse 3-D middle encoders on the hot path: SpMiddleFHD (rulebooks + sparse convs + dense() on the B200 kernels)."""
from .scn import SpMiddleFHD

__all__ = ["SpMiddleFHD"]
