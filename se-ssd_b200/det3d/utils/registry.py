"""String -> class registries (reference: det3d/utils/registry.py:6-76)."""
import inspect


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __repr__(self):
        return "%s(name=%s, items=%s)" % (type(self).__name__, self._name, list(self._module_dict))

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, cls):
        """Class decorator; a second class of the same name is an error (KeyError), as in the reference (:36-39)."""
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got %s" % type(cls))
        if cls.__name__ in self._module_dict:
            raise KeyError("%s is already registered in %s" % (cls.__name__, self._name))
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """Pop ``type`` from a config dict, look it up (KeyError when unknown) and call it with the remaining keys
    plus ``default_args`` for keys the config does not set (reference :47-76)."""
    if not (isinstance(cfg, dict) and "type" in cfg):
        raise AssertionError("cfg must be a dict with a 'type' key")
    if not (default_args is None or isinstance(default_args, dict)):
        raise AssertionError("default_args must be a dict or None")
    kwargs = dict(cfg)
    kind = kwargs.pop("type")
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError("%s is not in the %s registry" % (kind, registry.name))
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError("type must be a str or valid type, but got %s" % type(kind))
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return cls(**kwargs)
