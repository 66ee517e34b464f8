"""Import alias: ``import ist_b200`` == the package in ``can-image-style-transfer-save-automotive-radar_b200/``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("can-image-style-transfer-save-automotive-radar_b200")
sys.modules[__name__] = _pkg
