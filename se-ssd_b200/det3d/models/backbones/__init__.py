from .scn import SpMiddleFHD

__all__ = ["SpMiddleFHD"]
