"""B200-native Gatys style-transfer inner loop (the IST/ hot path of DJNing/Can-Image-Style-Transfer-Save-Automotive-Radar).

The directory name carries hyphens (it is the name the build contract fixes); import it as ``ist_b200`` through the
alias module at the repository root, or with ``importlib.import_module("can-image-style-transfer-save-automotive-radar_b200")``.
Compute goes through ``libist_b200.so`` (hand-written sm_100a CUDA behind the C ABI of include/ist_b200.h); there is no
CPU or PyTorch fallback.
"""
from . import _lib
from ._lib import IstError, build, load
from .plan import Plan, layer_table

__all__ = ["_lib", "IstError", "build", "load", "Plan", "layer_table"]
