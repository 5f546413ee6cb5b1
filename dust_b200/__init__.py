"""dust_b200 — B200-native (sm_100a) replacement for the ray-traversal hot path of dust-engine/dust.

Layout:  csrc/ (CUDA kernels + C ABI + C++ host mirror), lib/ (built .so files, in-tree),
capi.py (ctypes binding of include/dust_b200.h), host.py (binding of the host mirror),
pipeline.py / loader.py (object wrappers), scenes.py (seeded stand-in assets + .vox writer).
Importing the package does not load CUDA; `capi.load()` / `Context()` do and fail loudly.
"""
from . import capi  # noqa: F401

__all__ = ["capi", "host", "pipeline", "loader", "scenes"]
