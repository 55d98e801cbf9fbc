"""Python-file configs with attribute access (reference: det3d/torchie/utils/config.py:12-160).  The reference builds on the
third-party ``addict`` package (not installed here); ``ConfigDict`` below is a small self-contained replacement with the same
observable behaviour: nested dicts become ConfigDicts, missing keys raise KeyError / AttributeError."""
import os.path as osp
import sys
from importlib import import_module


class ConfigDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __missing__(self, name):
        raise KeyError(name)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    def __setattr__(self, name, value):
        self[name] = value

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}


class Config(object):
    @staticmethod
    def fromfile(filename):
        filename = osp.abspath(osp.expanduser(filename))
        if not osp.isfile(filename):
            raise FileNotFoundError('file "%s" does not exist' % filename)
        if not filename.endswith(".py"):
            raise IOError("Only py type configs are supported")
        name = osp.basename(filename)[:-3]
        if "." in name:
            raise ValueError("Dots are not allowed in config file path.")
        sys.path.insert(0, osp.dirname(filename))
        try:
            sys.modules.pop(name, None)
            mod = import_module(name)
        finally:
            sys.path.pop(0)
        cfg = {k: v for k, v in vars(mod).items() if not k.startswith("__")}
        return Config(cfg, filename=filename)

    def __init__(self, cfg_dict=None, filename=None):
        if cfg_dict is None:
            cfg_dict = {}
        elif not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got %s" % type(cfg_dict))
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        text = ""
        if filename:
            with open(filename, "r") as f:
                text = f.read()
        object.__setattr__(self, "_text", text)

    filename = property(lambda self: self._filename)
    text = property(lambda self: self._text)

    def __repr__(self):
        return "Config (path: %s): %r" % (self._filename, self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)
