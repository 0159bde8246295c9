"""corona-13 hot path, B200-native: Python side (tests / bench / tooling only).

The product is `libcorona_b200.so` (CUDA, sm_100a) behind the C ABI in include/corona_b200.h and the
plain-C host layer in host/ that mirrors the reference's accel.h module API.  This package holds
numpy record views, procedural scenes and the ctypes binding used by tests and bench.py.
The directory name contains a hyphen: import with importlib.import_module("corona-13_b200").
"""
from . import records, scenes, scene_io  # noqa: F401
