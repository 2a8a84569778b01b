"""hot_b200 — B200-native implicit-MPM hot path of penn-graphics-research/HOT.

The product is the CUDA library ``hot_b200/lib/libhot_b200.so`` (C ABI: ``include/hot_b200.h``).
This package is the thin host-side mirror used by tests and bench.py; it never falls back to a CPU path:
if the CUDA library is missing or no GPU is usable, construction raises.
"""
from ._lib import load_library, LibraryMissing  # noqa: F401
from .sim import MpmSimulationB200, HotError  # noqa: F401
