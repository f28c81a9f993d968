"""Importable alias: the package directory name (`3d-object-detection-for-autonomous-navigation_b200`)
is not a Python identifier, so `import pillars_b200` loads it through importlib."""
import importlib
import sys

_pkg = importlib.import_module("3d-object-detection-for-autonomous-navigation_b200")
sys.modules[__name__] = _pkg
