"""Minimal stand-in for `gymnasium` (TEST INFRASTRUCTURE ONLY, not product code).

`gymnasium` is not installed in the build container, and the reference NumPy env
(`/root/reference/tetris_gymnasium/envs/tetris.py`) imports it.  This package provides
just the names the reference touches so that the UNMODIFIED reference can be imported
to (a) validate the C oracle in `oracle/` and (b) generate the golden fixtures under
`tests/golden/` (see `oracle/make_golden.py`).  It is put on `sys.path` only by
`oracle/_refload.py`; nothing in the product package imports it.
"""
from . import core, spaces, utils  # noqa: F401
from .core import Env, ObservationWrapper, Wrapper  # noqa: F401

__version__ = "0.0-shim"

_REGISTRY = {}


def register(id, entry_point, **kwargs):  # noqa: A002
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):  # noqa: A002
    import importlib

    entry_point, base_kwargs = _REGISTRY[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod, name = entry_point.split(":")
        cls = getattr(importlib.import_module(mod), name)
    kw = dict(base_kwargs)
    kw.update(kwargs)
    return cls(**kw)
