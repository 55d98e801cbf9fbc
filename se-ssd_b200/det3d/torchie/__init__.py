from .utils import Config, ConfigDict


def is_str(x):
    return isinstance(x, str)


__all__ = ["Config", "ConfigDict", "is_str"]
