"""Importable alias for the package directory `high-fidelity-pointcloud-fusion_b200/` (its name has hyphens)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("high-fidelity-pointcloud-fusion_b200")
sys.modules[__name__] = _pkg
