"""Import-name drop-in: ``import polyblur`` served by this package.

The reference installs as ``polyblur`` (setup.py) with the modules ``deblurring``, ``blur_estimation``, ``filters``,
``edgetaper``, ``domain_transform`` and ``utils`` (polyblur/__init__.py:1 re-exports ``polyblur_deblurring`` and
``PolyblurDeblurring``).  An application that cannot edit its imports calls :func:`install` once, before its first
``import polyblur``; after that ``from polyblur.deblurring import inverse_filtering_rank3`` and friends resolve to the
CUDA engine.  Nothing is registered implicitly: a process that has the real reference on its path (``bench.py --impl
reference`` does) keeps it.
"""
from __future__ import annotations

import importlib
import sys

_SUBMODULES = ("deblurring", "blur_estimation", "filters", "edgetaper", "domain_transform", "utils")


def install(name: str = "polyblur", force: bool = False) -> None:
    """Registers this package and its reference-named submodules in ``sys.modules`` under ``name``.

    Raises ``ImportError`` when a different package of that name is already imported (``force=True`` replaces it)."""
    pkg = importlib.import_module(__package__)
    have = sys.modules.get(name)
    if have is not None and have is not pkg and not force:
        raise ImportError(f"a different '{name}' is already imported from {getattr(have, '__file__', '?')}; "
                          "call install() before the first import, or pass force=True")
    if force:
        for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            del sys.modules[key]
    sys.modules[name] = pkg
    for sub in _SUBMODULES:
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__package__}.{sub}")


def uninstall(name: str = "polyblur") -> None:
    """Removes what :func:`install` registered (only entries that point at this package)."""
    pkg = importlib.import_module(__package__)
    for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
        mod = sys.modules[key]
        if mod is pkg or getattr(mod, "__package__", "") == __package__ or getattr(mod, "__name__", "").startswith(__package__ + "."):
            del sys.modules[key]
