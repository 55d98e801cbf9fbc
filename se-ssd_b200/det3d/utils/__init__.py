"""Registry / config helpers the reference's config file and builders import from ``det3d.utils``."""
from .registry import Registry, build_from_cfg

__all__ = ["Registry", "build_from_cfg"]
