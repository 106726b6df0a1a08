"""gym.envs.registration stand-in (envs/__init__.py:1-5 of the reference)."""
import importlib

_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = _registry[id]
    mod_name, cls_name = entry_point.split(':')
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(kw.get('kwargs', {}))
    kw.update(kwargs)
    return cls(**kw)
