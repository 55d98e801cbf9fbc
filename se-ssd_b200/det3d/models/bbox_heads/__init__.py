from .mg_head_sessd import Head, MultiGroupHead

__all__ = ["Head", "MultiGroupHead"]
