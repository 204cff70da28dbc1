"""Hand names this tree does not define through to the reference's module of the same name.

The reference's `innovative3D/` directory (a namespace portion: no `__init__.py`) is on this package's
`__path__` after this tree's own directory when a reference checkout is on `sys.path`
(`__init__.py`). `reference_module("config")` loads `<that dir>/config.py` under the private name
`innovative3D._reference_config` (never under the public name: `innovative3D.config` stays this tree's),
and `module_getattr` is the PEP 562 `__getattr__` hook `config.py` / `helpers.py` / `models.py` install."""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path
from types import ModuleType
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LOADING: set = set()


def reference_dirs():
    """Other `innovative3D` directories on this package's __path__ (the reference checkout, if any)."""
    import innovative3D
    return [p for p in innovative3D.__path__ if os.path.abspath(p) != _HERE and os.path.isdir(p)]


def reference_module(name: str) -> Optional[ModuleType]:
    """The reference's `innovative3D/<name>.py`, imported once as `innovative3D._reference_<name>`, or None."""
    key = f"innovative3D._reference_{name}"
    if key in sys.modules:
        return sys.modules[key]
    if key in _LOADING:      # the reference module imports this tree's module, which asks again: not ready yet
        return None
    for d in reference_dirs():
        path = os.path.join(d, name + ".py")
        if not os.path.isfile(path):
            continue
        spec = importlib.util.spec_from_file_location(key, path)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "innovative3D"
        sys.modules[key] = mod
        _LOADING.add(key)
        _mkdir = Path.mkdir

        def _safe_mkdir(self, *a, **k):   # config.py:15-19 creates directories under a hard-coded home path
            try:
                return _mkdir(self, *a, **k)
            except OSError:
                return None

        Path.mkdir = _safe_mkdir
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            sys.modules.pop(key, None)
            raise
        finally:
            Path.mkdir = _mkdir
            _LOADING.discard(key)
        return mod
    return None


def module_getattr(module_name: str, short: str, what: str):
    """PEP 562 hook for `innovative3D.<short>`: look the missing attribute up in the reference's module."""
    def __getattr__(attr: str):
        if attr.startswith("__"):
            raise AttributeError(attr)
        ref = reference_module(short)
        if ref is not None and hasattr(ref, attr):
            return getattr(ref, attr)
        where = "the reference checkout on sys.path does not define it either" if ref is not None else \
            "no reference checkout is on sys.path to provide it"
        raise AttributeError(f"module '{module_name}' (B200 hot-path build) has no attribute '{attr}': {what}; {where}")
    return __getattr__
