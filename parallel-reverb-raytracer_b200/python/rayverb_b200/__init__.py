"""rayverb_b200: host-side Python plumbing over the C ABI (include/rayverb_b200.h).

The product is the CUDA library (csrc/) and the C++ mirror of rayverb's classes (host/);
this package only holds the ctypes binding used by tests/ and bench.py plus the synthetic
scene generators for the BASELINE.json configurations.
"""
from . import scenes, types  # noqa: F401
