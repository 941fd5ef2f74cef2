"""gsgp-b200: B200-native (sm_100a) generation-loop hot path of go-gsgp behind a C-ABI.

csrc/  CUDA kernels + the C-ABI (include/gsgp_b200.h) -> libgsgp_b200.so
capi   ctypes binding of the C-ABI
engine host-side mirror of the reference's hot-path functions
"""
from . import capi  # noqa: F401
from .engine import Engine, GsgpError, pack_trees  # noqa: F401

__all__ = ["Engine", "GsgpError", "pack_trees", "capi"]
