"""ctypes binding of the C-ABI in include/gsgp_b200.h (libgsgp_b200.so).

There is no CPU fallback: if the CUDA library is missing or no device is present, every entry
point raises.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgsgp_b200.so")

MSE, RMSE, MAE, MRE = 0, 1, 2, 3
OP_INIT, OP_XO, OP_MUT, OP_REPR = 0, 1, 2, 3
RNG_HOST_REPLAY, RNG_DEVICE_PHILOX = 0, 1
FLAG_KEEP_TREE_SEMANTICS = 1
FLAG_LINEAR_SCALING = 2
FLAG_KEEP_RUN_HISTORY = 4
FLAG_SINGLE_BUFFER_STORE = 8
K_INTERP, K_GSOP, K_FINALIZE, K_PLAN = 0, 1, 2, 3


class GsgpConfig(C.Structure):
    _fields_ = [
        ("population_size", C.c_int32), ("nvar", C.c_int32),
        ("nrow", C.c_int32), ("nrow_test", C.c_int32),
        ("max_depth_creation", C.c_int32), ("tournament_size", C.c_int32),
        ("zero_depth", C.c_int32), ("minimization_problem", C.c_int32),
        ("error_measure", C.c_int32), ("num_constants", C.c_int32),
        ("p_crossover", C.c_double), ("p_mutation", C.c_double),
        ("rng_seed", C.c_int64), ("rng_mode", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("nccl_id", C.c_uint8 * 128),
        ("workspace_bytes", C.c_int64), ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


PLAN_DTYPE = np.dtype(
    [("op", "<i4"), ("elite", "<i4"), ("p1", "<i4"), ("p2", "<i4"),
     ("tree0_off", "<i4"), ("tree0_len", "<i4"), ("tree1_off", "<i4"), ("tree1_len", "<i4"),
     ("ms", "<f8")], align=True)
assert PLAN_DTYPE.itemsize == 40

# every symbol include/gsgp_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p
SIGNATURES = {
    "gsgp_abi_version": (C.c_int32, []),
    "gsgp_config_defaults": (None, [C.POINTER(GsgpConfig)]),
    "gsgp_last_error": (C.c_char_p, [_vp]),
    "gsgp_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "gsgp_shard_range": (None, [C.c_int32, C.c_int32, C.c_int32, _ip, _ip]),
    "gsgp_create": (C.c_int, [C.POINTER(GsgpConfig), C.POINTER(_vp)]),
    "gsgp_destroy": (None, [_vp]),
    "gsgp_local_counts": (C.c_int, [_vp, _ip, _ip]),
    "gsgp_store_info": (C.c_int, [_vp, _ip, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gsgp_exchange_mode": (C.c_int, [_vp]),
    "gsgp_set_dataset": (C.c_int, [_vp, _dp]),
    "gsgp_set_dataset_shard": (C.c_int, [_vp, _dp, _dp]),
    "gsgp_set_constants": (C.c_int, [_vp, _dp, C.c_int32]),
    "gsgp_seed_semantic": (C.c_int, [_vp, C.c_int32, _dp]),
    "gsgp_seed_semantic_files": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_char_p)]),
    "gsgp_eval_initial": (C.c_int, [_vp, C.c_int32, C.c_int32, _ip, C.c_int32, _ip, _ip]),
    "gsgp_fitness_all": (C.c_int, [_vp, _dp, _dp, _ip]),
    "gsgp_generation": (C.c_int, [_vp, _vp, _ip, C.c_int32, _dp, _dp, _ip]),
    "gsgp_generation_compiled": (C.c_int, [_vp, _vp, _ip, C.c_int32, _vp, C.c_int32, _ip, _ip, _dp, _dp, _ip]),
    "gsgp_run_generations": (C.c_int, [_vp, C.c_int32]),
    "gsgp_sync": (C.c_int, [_vp]),
    "gsgp_get_fitness": (C.c_int, [_vp, _dp, _dp, _ip]),
    "gsgp_get_semantic": (C.c_int, [_vp, C.c_int32, _dp]),
    "gsgp_get_tree_semantic": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp]),
    "gsgp_get_plan": (C.c_int, [_vp, _vp]),
    "gsgp_get_tree": (C.c_int, [_vp, C.c_int32, C.c_int32, _ip, C.c_int32, _ip]),
    "gsgp_get_history": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp, _dp, _ip]),
    "gsgp_get_best_semantic_of": (C.c_int, [_vp, C.c_int32, _dp]),
    "gsgp_get_plan_of": (C.c_int, [_vp, C.c_int32, _vp]),
    "gsgp_generation_index": (C.c_int32, [_vp]),
    "gsgp_eval_trees": (C.c_int, [_vp, C.c_int32, _ip, C.c_int32, _ip, _ip, _dp]),
    "gsgp_set_fitness": (C.c_int, [_vp, _dp, _dp]),
    "gsgp_draw_plan": (C.c_int, [_vp, C.c_int32, _vp]),
    "gsgp_stream": (_vp, [_vp]),
    "gsgp_profile_enable": (C.c_int, [_vp, C.c_int32]),
    "gsgp_profile_read": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_int64), _dp, _dp]),
    "gsgp_launch_count": (C.c_int64, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libgsgp_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m go_gsgp_b200.build` "
            "(the gsgp_b200 path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.gsgp_abi_version() != 1:
        raise RuntimeError("libgsgp_b200.so ABI version mismatch")
    _lib = L
    return L


def dptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)
