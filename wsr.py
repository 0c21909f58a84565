"""Import shim: the product package directory carries the (hyphenated) name of the
reference repository, which is not a Python identifier.  ``import wsr`` loads it with
importlib and re-exports it, so callers write ``wsr.pkg`` / ``wsr.native`` etc.

Always reach sub-modules by attribute access on ``wsr.pkg`` (never ``import wsr.x``),
so every module exists exactly once in ``sys.modules``.
"""
import importlib
import os
import sys

PKG_NAME = "super-resolution-enhancement-of-weather-data-using-diffusion-models_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module(PKG_NAME)


def sub(name: str):
    """Return sub-module ``name`` (dotted, relative to the package), importing it once."""
    return importlib.import_module(PKG_NAME + "." + name)
