from . import pipelines  # noqa: F401
from .registry import DATASETS, PIPELINES

__all__ = ["DATASETS", "PIPELINES"]
