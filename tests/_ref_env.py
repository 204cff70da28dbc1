"""Test helper: run the reference's own callers (`train.py`, `test.py`) on top of this repo's `innovative3D` package.

Where the reference comes from: `/root/reference` in the build container, or the git-ignored staging copy
`baseline/_ref/` that `__graft_entry__.build()` makes and `gpurun` ships to the GPU box. Neither present -> the tests
that need it skip. Third-party packages the reference imports for plotting / file IO and that this image lacks
(matplotlib, seaborn, pydicom, torchmetrics, thop) are replaced by inert stubs; `pytorch_lightning` by the repo's
stand-in (`innovative3D._lightning.install()`) unless the real one is installed. None of the reference's arithmetic is
stubbed."""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "spff-unet-spcct_b200"


def reference_root():
    for p in (Path("/root/reference"), ROOT / "baseline" / "_ref"):
        if (p / "train.py").is_file() and (p / "innovative3D" / "models.py").is_file():
            return p
    return None


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Any()

    def __setitem__(self, k, v):
        pass


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        if importlib.util.find_spec(name) is not None:
            return None          # the real package exists: leave it alone
    except (ImportError, ValueError):
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _missing(k):                     # any other attribute: an inert object
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any()

    m.__getattr__ = _missing
    sys.modules[name] = m
    return m


def prepare(tmp_dir) -> Path:
    """sys.path / sys.modules / environment so that `import train` (the reference's script) resolves `innovative3D` to this
    repo's package, with the reference checkout behind it. Returns the reference root (pytest.skip when there is none)."""
    import pytest
    ref = reference_root()
    if ref is None:
        pytest.skip("no reference checkout (/root/reference or baseline/_ref) on this machine")
    os.environ["CHECKPOINT_DIR"] = str(Path(tmp_dir) / "ckpt")
    os.environ["LOG_DIR"] = str(Path(tmp_dir) / "logs")
    os.environ.setdefault("FAST_SKIP_VIZ", "1")
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    if mpl is not None:
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.patches = _stub("matplotlib.patches", Patch=_Any)
        _stub("matplotlib.colors")
        _stub("matplotlib.cm")
    for name in ("seaborn", "pydicom", "torchmetrics", "statsmodels"):
        _stub(name)
    for p in (str(ROOT), str(PKG)):
        if p not in sys.path:
            sys.path.insert(0, p)
    if str(ref) not in sys.path:
        sys.path.append(str(ref))          # behind this repo's package; a regular package wins over it anyway
    # a package imported before the reference was on the path has a stale __path__: extend it again
    if "innovative3D" in sys.modules:
        from pkgutil import extend_path
        pkg = sys.modules["innovative3D"]
        pkg.__path__ = extend_path(list(pkg.__path__), pkg.__name__)
    from innovative3D import _lightning
    _lightning.install()
    return ref


def import_script(ref: Path, name: str):
    """Import the reference's top-level script `<name>.py` as module `ref_<name>` (not `__main__`: nothing runs)."""
    key = f"ref_{name}"
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, ref / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
