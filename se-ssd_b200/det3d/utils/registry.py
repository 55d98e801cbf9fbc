"""String -> class registries with the reference's surface (det3d/utils/registry.py): ``Registry(name)``, ``.name``, ``.module_dict``,
``.get(key)``, ``@registry.register_module`` (duplicate names are a KeyError, non-classes a TypeError) and
``build_from_cfg(cfg, registry, default_args)`` (pops ``type``; unknown type strings are a KeyError; ``default_args`` only fill keys the
config leaves unset)."""
import inspect


class Registry(object):
    def __init__(self, name):
        self._name, self._module_dict = name, {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __repr__(self):
        return "{}(name={}, items={})".format(type(self).__name__, self._name, sorted(self._module_dict))

    def __contains__(self, key):
        return key in self._module_dict

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got %s" % type(cls))
        key = cls.__name__
        if key in self:
            raise KeyError("%s is already registered in %s" % (key, self._name))
        self._module_dict[key] = cls
        return cls


def _resolve(kind, registry):
    if inspect.isclass(kind):
        return kind
    if not isinstance(kind, str):
        raise TypeError("type must be a str or valid type, but got %s" % type(kind))
    found = registry.get(kind)
    if found is None:
        raise KeyError("%s is not in the %s registry" % (kind, registry.name))
    return found


def build_from_cfg(cfg, registry, default_args=None):
    assert isinstance(cfg, dict) and "type" in cfg, "cfg must be a dict with a 'type' key"
    assert default_args is None or isinstance(default_args, dict), "default_args must be a dict or None"
    kwargs = {k: v for k, v in cfg.items() if k != "type"}
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return _resolve(cfg["type"], registry)(**kwargs)
