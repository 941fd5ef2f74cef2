"""ctypes binding of the CPU oracle (oracle/gsgp_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package (go_gsgp_b200) never imports this.
PARITY UNPINNED by the reference's own tests (see gsgp_oracle.h); the random stream is Go's own, pinned by known
outputs of Go programs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgsgp_oracle.so")

MSE, RMSE, MAE, MRE = 0, 1, 2, 3
OP_INIT, OP_XO, OP_MUT, OP_REPR = 0, 1, 2, 3
RNG_ALFG, RNG_PHILOX = 0, 1
EXP_LIBM, EXP_GO_PORTABLE = 0, 1


class OrcConfig(C.Structure):
    _fields_ = [
        ("population_size", C.c_int32),
        ("init_type", C.c_int32),
        ("p_crossover", C.c_double),
        ("p_mutation", C.c_double),
        ("max_depth_creation", C.c_int32),
        ("tournament_size", C.c_int32),
        ("zero_depth", C.c_int32),
        ("num_random_constants", C.c_int32),
        ("min_random_constant", C.c_double),
        ("max_random_constant", C.c_double),
        ("minimization_problem", C.c_int32),
        ("error_measure", C.c_int32),
        ("n_workers", C.c_int32),
        ("rng_kind", C.c_int32),
        ("rng_seed", C.c_int64),
        ("use_linear_scaling", C.c_int32),
        ("reserved", C.c_int32),
    ]


class OrcPlanEntry(C.Structure):
    _fields_ = [
        ("op", C.c_int32),
        ("elite", C.c_int32),
        ("p1", C.c_int32),
        ("p2", C.c_int32),
        ("tree0_off", C.c_int32),
        ("tree0_len", C.c_int32),
        ("tree1_off", C.c_int32),
        ("tree1_len", C.c_int32),
        ("ms", C.c_double),
    ]


PLAN_DTYPE = np.dtype(
    [("op", "<i4"), ("elite", "<i4"), ("p1", "<i4"), ("p2", "<i4"),
     ("tree0_off", "<i4"), ("tree0_len", "<i4"), ("tree1_off", "<i4"), ("tree1_len", "<i4"),
     ("ms", "<f8")], align=True)
assert PLAN_DTYPE.itemsize == C.sizeof(OrcPlanEntry)


class OrcRng(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("vec", C.c_uint64 * 607),
        ("tap", C.c_int),
        ("feed", C.c_int),
        ("key", C.c_uint32 * 2),
        ("stream", C.c_uint32 * 3),
        ("block", C.c_uint32),
        ("out", C.c_uint32 * 4),
        ("have", C.c_int),
        ("n_draws", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("gsgp_oracle.c", "gsgp_oracle.h", "oracle_cli.c", "go_rng_cooked.inc.h", "Makefile")
        if os.path.exists(os.path.join(_HERE, f))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    sig = {
        "orc_config_defaults": (None, [C.POINTER(OrcConfig)]),
        "orc_new": (vp, [C.POINTER(OrcConfig), C.c_int32, C.c_int32, C.c_int32, dp]),
        "orc_free": (None, [vp]),
        "orc_get_rng": (C.POINTER(OrcRng), [vp]),
        "orc_create_T_F": (None, [vp]),
        "orc_seed_semantic": (C.c_int, [vp, C.c_int32, dp]),
        "orc_initialize_population": (None, [vp, C.c_int32]),
        "orc_evaluate": (None, [vp]),
        "orc_set_initial_tree": (None, [vp, C.c_int32, ip, C.c_int32]),
        "orc_generation": (None, [vp]),
        "orc_generation_replay": (None, [vp, C.c_void_p, ip]),
        "orc_effects_step": (None, [vp, C.c_int]),
        "orc_set_exp_kind": (None, [C.c_int]),
        "orc_exp64": (C.c_double, [C.c_double]),
        "orc_sigmoid": (C.c_double, [C.c_double]),
        "orc_create_grow_tree_arrays": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32, ip, C.c_int32]),
        "orc_eval_arrays": (C.c_double, [vp, ip, C.c_int32, C.c_int32]),
        "orc_semantic_evaluate_array": (None, [vp, ip, C.c_int32, C.c_int32, dp]),
        "orc_fitness_train": (C.c_double, [vp, dp]),
        "orc_fitness_test": (C.c_double, [vp, dp]),
        "orc_fitness_train_ls": (C.c_double, [vp, dp, dp, dp]),
        "orc_fitness_test_ls": (C.c_double, [vp, dp, C.c_double, C.c_double]),
        "orc_tournament_selection": (C.c_int32, [vp]),
        "orc_best_individual": (C.c_int32, [vp]),
        "orc_better": (C.c_int, [vp, C.c_double, C.c_double]),
        "orc_gs_crossover_vec": (None, [dp, dp, dp, dp, C.c_int64]),
        "orc_gs_mutation_vec": (None, [dp, dp, dp, C.c_double, dp, C.c_int64]),
        "orc_generation_index": (C.c_int32, [vp]),
        "orc_index_best": (C.c_int32, [vp]),
        "orc_fit": (dp, [vp]),
        "orc_fit_test": (dp, [vp]),
        "orc_sem_train": (dp, [vp, C.c_int32]),
        "orc_sem_test": (dp, [vp, C.c_int32]),
        "orc_last_plan": (C.c_void_p, [vp]),
        "orc_last_tree_pool": (ip, [vp, ip]),
        "orc_initial_tree": (ip, [vp, C.c_int32, ip]),
        "orc_symbol_values": (dp, [vp, ip]),
        "orc_last_tree_semantic": (dp, [vp, C.c_int32, C.c_int32]),
        "orc_keep_tree_semantics": (None, [vp, C.c_int]),
        "orc_format_float": (C.c_int, [C.c_double, C.c_char_p, C.c_size_t]),
        "orc_read_dataset": (C.c_int, [C.c_char_p, C.c_char_p, ip, ip, ip, C.POINTER(dp)]),
        "orc_read_sem": (C.c_int, [C.c_char_p, C.c_int32, dp]),
        "orc_rng_seed": (None, [C.POINTER(OrcRng), C.c_int, C.c_int64]),
        "orc_rng_seek": (None, [C.POINTER(OrcRng), C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_rng_uint64": (C.c_uint64, [C.POINTER(OrcRng)]),
        "orc_rng_int63": (C.c_int64, [C.POINTER(OrcRng)]),
        "orc_rng_int31": (C.c_int32, [C.POINTER(OrcRng)]),
        "orc_rng_float64": (C.c_double, [C.POINTER(OrcRng)]),
        "orc_rng_intn": (C.c_int32, [C.POINTER(OrcRng), C.c_int32]),
        "orc_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def make_config(**kw) -> OrcConfig:
    c = OrcConfig()
    lib().orc_config_defaults(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


class Oracle:
    """Thin object wrapper over orc_state, named after the reference's functions."""

    def __init__(self, cfg: OrcConfig, X: np.ndarray, y: np.ndarray, nrow: int):
        """X: (N, V) float64, y: (N,), first nrow rows are the train set (main_cpu.go:406-420)."""
        self.L = lib()
        self.cfg = cfg
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        self.N, self.nvar = X.shape
        self.nrow = int(nrow)
        self.nrow_test = self.N - self.nrow
        self.rowmajor = np.ascontiguousarray(np.concatenate([X, y[:, None]], axis=1))
        self.h = self.L.orc_new(C.byref(cfg), self.nvar, self.nrow, self.nrow_test, _dp(self.rowmajor))
        self.P = cfg.population_size

    def close(self):
        if self.h:
            self.L.orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- setup -------------------------------------------------------------------------
    def create_T_F(self):
        self.L.orc_create_T_F(self.h)

    def seed_semantic(self, ind: int, sem: np.ndarray):
        sem = np.ascontiguousarray(sem, dtype=np.float64)
        assert sem.shape == (self.N,)
        assert self.L.orc_seed_semantic(self.h, ind, _dp(sem)) == 0

    def initialize_population(self, n_seeds: int = 0):
        self.L.orc_initialize_population(self.h, n_seeds)

    def set_initial_tree(self, ind: int, tree: np.ndarray):
        tree = np.ascontiguousarray(tree, dtype=np.int32)
        self.L.orc_set_initial_tree(self.h, ind, _ip(tree), len(tree))

    def evaluate(self):
        self.L.orc_evaluate(self.h)

    def keep_tree_semantics(self, flag=True):
        self.L.orc_keep_tree_semantics(self.h, int(flag))

    # --- generation loop ---------------------------------------------------------------
    def generation(self):
        self.L.orc_generation(self.h)

    def effects_step(self, test_crossover: bool):
        """operators_effects.go's single step (:783-806,836-842,1246-1256); last_plan() / last_tree_pool() hold its draws"""
        self.L.orc_effects_step(self.h, int(bool(test_crossover)))

    def generation_replay(self, plan: np.ndarray, pool: np.ndarray):
        plan = np.ascontiguousarray(plan)
        assert plan.dtype == PLAN_DTYPE and len(plan) == self.P
        pool = np.ascontiguousarray(pool, dtype=np.int32)
        if len(pool) == 0:
            pool = np.zeros(1, np.int32)
        self.L.orc_generation_replay(self.h, plan.ctypes.data, _ip(pool))

    # --- accessors ---------------------------------------------------------------------
    @property
    def index_best(self) -> int:
        return self.L.orc_index_best(self.h)

    @property
    def generation_index(self) -> int:
        return self.L.orc_generation_index(self.h)

    def fit(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_fit(self.h), (self.P,)).copy()

    def fit_test(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_fit_test(self.h), (self.P,)).copy()

    def sem_train(self, ind: int) -> np.ndarray:
        if self.nrow == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(self.L.orc_sem_train(self.h, ind), (self.nrow,)).copy()

    def sem_test(self, ind: int) -> np.ndarray:
        if self.nrow_test == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(self.L.orc_sem_test(self.h, ind), (self.nrow_test,)).copy()

    def semantic(self, ind: int) -> np.ndarray:
        return np.concatenate([self.sem_train(ind), self.sem_test(ind)])

    def semantics(self) -> np.ndarray:
        return np.stack([self.semantic(i) for i in range(self.P)])

    def last_plan(self) -> np.ndarray:
        buf = (C.c_char * (PLAN_DTYPE.itemsize * self.P)).from_address(self.L.orc_last_plan(self.h))
        return np.frombuffer(buf, dtype=PLAN_DTYPE).copy()

    def last_tree_pool(self) -> np.ndarray:
        n = C.c_int32(0)
        p = self.L.orc_last_tree_pool(self.h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def initial_tree(self, ind: int):
        n = C.c_int32(0)
        p = self.L.orc_initial_tree(self.h, ind, C.byref(n))
        if not p or n.value == 0:
            return None
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def symbol_values(self) -> np.ndarray:
        n = C.c_int32(0)
        p = self.L.orc_symbol_values(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def last_tree_semantic(self, ind: int, which: int):
        p = self.L.orc_last_tree_semantic(self.h, ind, which)
        if not p:
            return None
        return np.ctypeslib.as_array(p, (self.N,)).copy()

    # --- primitives --------------------------------------------------------------------
    def create_grow_tree_arrays(self, max_depth: int | None = None) -> np.ndarray:
        d = self.cfg.max_depth_creation if max_depth is None else max_depth
        cap = 4 * (1 << max(d, 1)) + 8
        out = np.zeros(cap, np.int32)
        n = self.L.orc_create_grow_tree_arrays(self.h, 0, d, 0, _ip(out), cap)
        return out[:n].copy()

    def semantic_evaluate_array(self, tree: np.ndarray) -> np.ndarray:
        """train then test semantics of one tree (main_cpu.go:889-890)."""
        tree = np.ascontiguousarray(tree, dtype=np.int32)
        out = np.zeros(self.N)
        tr, te = out[: self.nrow], out[self.nrow:]
        if self.nrow:
            self.L.orc_semantic_evaluate_array(self.h, _ip(tree), self.nrow, 0, _dp(tr))
        if self.nrow_test:
            self.L.orc_semantic_evaluate_array(self.h, _ip(tree), self.nrow_test, self.nrow, _dp(te))
        return out

    def fitness_train(self, sem_train: np.ndarray) -> float:
        return self.L.orc_fitness_train(self.h, _dp(np.ascontiguousarray(sem_train, dtype=np.float64)))

    def fitness_test(self, sem_test: np.ndarray) -> float:
        return self.L.orc_fitness_test(self.h, _dp(np.ascontiguousarray(sem_test, dtype=np.float64)))

    def fitness_train_ls(self, sem_train: np.ndarray):
        """-> (fitness, a, b)  (fitness_of_semantic_train_ls main_cpu.go:991-1104)"""
        a, b = C.c_double(), C.c_double()
        f = self.L.orc_fitness_train_ls(self.h, _dp(np.ascontiguousarray(sem_train, dtype=np.float64)),
                                        C.cast(C.byref(a), C.POINTER(C.c_double)), C.cast(C.byref(b), C.POINTER(C.c_double)))
        return f, a.value, b.value

    def fitness_test_ls(self, sem_test: np.ndarray, a: float, b: float) -> float:
        return self.L.orc_fitness_test_ls(self.h, _dp(np.ascontiguousarray(sem_test, dtype=np.float64)), a, b)

    def tournament_selection(self) -> int:
        return self.L.orc_tournament_selection(self.h)

    def rng(self):
        return self.L.orc_get_rng(self.h)


def gs_crossover(p1, p2, r) -> np.ndarray:
    p1, p2, r = (np.ascontiguousarray(a, dtype=np.float64) for a in (p1, p2, r))
    out = np.empty_like(p1)
    lib().orc_gs_crossover_vec(_dp(p1), _dp(p2), _dp(r), _dp(out), p1.size)
    return out


def gs_mutation(t, r1, r2, ms) -> np.ndarray:
    t, r1, r2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, r1, r2))
    out = np.empty_like(t)
    lib().orc_gs_mutation_vec(_dp(t), _dp(r1), _dp(r2), float(ms), _dp(out), t.size)
    return out


def format_float(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_format_float(float(v), buf, 64)
    return buf.value.decode()


class Rng:
    """Stand-alone oracle RNG (for RNG KATs and plan-parity tests)."""

    def __init__(self, kind: int, seed: int):
        self.L = lib()
        self.r = OrcRng()
        self.L.orc_rng_seed(C.byref(self.r), kind, seed)

    def seek(self, s0, s1=0, s2=0):
        self.L.orc_rng_seek(C.byref(self.r), s0, s1, s2)

    def uint64(self):
        return self.L.orc_rng_uint64(C.byref(self.r))

    def float64(self):
        return self.L.orc_rng_float64(C.byref(self.r))

    def intn(self, n):
        return self.L.orc_rng_intn(C.byref(self.r), n)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)
