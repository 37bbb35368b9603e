"""Importable alias: ``import dreg_nerf_b200`` -> the package directory ``dreg-nerf_b200/``
(whose name, fixed by the project layout, is not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("dreg-nerf_b200")
sys.modules[__name__] = _pkg
