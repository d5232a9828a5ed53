"""lego_loam_b200 -- B200-native (sm_100a) drop-in for LeGO-LOAM's scan-matching hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/llb200.h), host/ (C++ adapter classes with
the reference's member-function names), api.py (ctypes mirror for tests / bench), synth.py
(synthetic range-image-shaped data).  The package never imports oracle/.
"""
from . import api, synth  # noqa: F401

__all__ = ["api", "synth"]
