"""Host-side mirror of the reference's hot-path functions over the C-ABI.

Method names follow main_cpu.go (init_tables, evaluate, update_tables happen inside the context):
the Go host keeps flags/ini, file readers, RNG draws (replay mode), contributions and output writers,
and calls these instead of its per-individual operators (see INTEGRATION.md for the cgo binding).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import PLAN_DTYPE, GsgpConfig


class GsgpError(RuntimeError):
    """The reference panics on every failure; the binding raises."""


def pack_trees(trees):
    """list of int32 arrays (None for absent) -> (pool, off, len)."""
    off = np.full(len(trees), -1, np.int32)
    ln = np.zeros(len(trees), np.int32)
    parts, pos = [], 0
    for i, t in enumerate(trees):
        if t is None or len(t) == 0:
            continue
        t = np.asarray(t, np.int32)
        off[i], ln[i] = pos, len(t)
        parts.append(t)
        pos += len(t)
    pool = np.concatenate(parts).astype(np.int32) if parts else np.zeros(1, np.int32)
    return np.ascontiguousarray(pool), off, ln


class Engine:
    def __init__(self, *, population_size, nvar, nrow, nrow_test, max_depth_creation=6,
                 tournament_size=4, zero_depth=False, minimization_problem=True, error_measure=capi.MSE,
                 num_constants=0, p_crossover=0.6, p_mutation=0.3, rng_seed=1,
                 rng_mode=capi.RNG_HOST_REPLAY, device=0, rank=0, world_size=1, nccl_id=None,
                 workspace_bytes=0, flags=0, use_linear_scaling=False):
        self.L = capi.load()
        cfg = GsgpConfig()
        self.L.gsgp_config_defaults(C.byref(cfg))
        cfg.population_size, cfg.nvar, cfg.nrow, cfg.nrow_test = population_size, nvar, nrow, nrow_test
        cfg.max_depth_creation, cfg.tournament_size = max_depth_creation, tournament_size
        cfg.zero_depth, cfg.minimization_problem = int(zero_depth), int(minimization_problem)
        cfg.error_measure, cfg.num_constants = error_measure, num_constants
        cfg.p_crossover, cfg.p_mutation, cfg.rng_seed = p_crossover, p_mutation, rng_seed
        cfg.rng_mode, cfg.device, cfg.rank, cfg.world_size = rng_mode, device, rank, world_size
        if use_linear_scaling:                                   # -linsc main_cpu.go:181
            flags |= capi.FLAG_LINEAR_SCALING
        cfg.workspace_bytes, cfg.flags = workspace_bytes, flags
        if nccl_id is not None:
            C.memmove(cfg.nccl_id, bytes(nccl_id), 128)
        self.cfg = cfg
        self.h = C.c_void_p()
        if self.L.gsgp_create(C.byref(cfg), C.byref(self.h)) != 0:
            raise GsgpError(self.L.gsgp_last_error(None).decode())
        self.P = population_size
        a, b = C.c_int32(), C.c_int32()
        self.L.gsgp_local_counts(self.h, C.byref(a), C.byref(b))
        self.nrow_local, self.nrow_test_local = a.value, b.value
        self.n_local = a.value + b.value
        self.n_symbols = 4 + nvar + num_constants
        nb, bb, sb = C.c_int32(), C.c_int64(), C.c_int64()
        self.L.gsgp_store_info(self.h, C.byref(nb), C.byref(bb), C.byref(sb))
        self.store_buffers, self.store_buffer_bytes, self.store_scratch_bytes = nb.value, bb.value, sb.value
        self.exchange = {0: "none", 1: "nccl", 2: "p2p"}[self.L.gsgp_exchange_mode(self.h)]

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise GsgpError(self.L.gsgp_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.gsgp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = capi.load()
        buf = (C.c_uint8 * 128)()
        if L.gsgp_nccl_unique_id(buf) != 0:
            raise GsgpError(L.gsgp_last_error(None).decode())
        return bytes(buf)

    @staticmethod
    def shard_range(n, rank, world):
        lo, hi = C.c_int32(), C.c_int32()
        capi.load().gsgp_shard_range(n, rank, world, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    # ------------------------------------------------------------------ inputs
    def set_dataset(self, X, y):
        """read_input_data's `set` (main_cpu.go:381-421): X (N,V), y (N,), train rows first."""
        rm = np.ascontiguousarray(np.concatenate([np.asarray(X, np.float64), np.asarray(y, np.float64)[:, None]], axis=1))
        assert rm.shape == (self.cfg.nrow + self.cfg.nrow_test, self.cfg.nvar + 1)
        self._ck(self.L.gsgp_set_dataset(self.h, capi.dptr(rm)))

    def set_dataset_shard(self, train_rows, test_rows):
        tr = np.ascontiguousarray(train_rows, np.float64)
        te = np.ascontiguousarray(test_rows, np.float64)
        self._ck(self.L.gsgp_set_dataset_shard(self.h, capi.dptr(tr), capi.dptr(te)))

    def set_constants(self, sym_val):
        sv = np.ascontiguousarray(sym_val, np.float64)
        self._ck(self.L.gsgp_set_constants(self.h, capi.dptr(sv), len(sv)))

    def seed_semantic(self, ind, sem):
        sem = np.ascontiguousarray(sem, np.float64)
        assert sem.shape == (self.cfg.nrow + self.cfg.nrow_test,)
        self._ck(self.L.gsgp_seed_semantic(self.h, ind, capi.dptr(sem)))

    def seed_semantic_files(self, first_ind, paths):
        """semantic feeding from the reference's positional files (main_cpu.go:1342-1349,1383-1388)."""
        arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
        self._ck(self.L.gsgp_seed_semantic_files(self.h, first_ind, len(paths), arr))

    # ------------------------------------------------------------------ evaluate()
    def eval_initial(self, first_ind, trees):
        pool, off, ln = pack_trees(trees)
        self._ck(self.L.gsgp_eval_initial(self.h, first_ind, len(trees), capi.iptr(pool), len(pool),
                                          capi.iptr(off), capi.iptr(ln)))

    def fitness_all(self):
        fit, fit_test, best = np.empty(self.P), np.empty(self.P), C.c_int32()
        self._ck(self.L.gsgp_fitness_all(self.h, capi.dptr(fit), capi.dptr(fit_test), C.byref(best)))
        return fit, fit_test, best.value

    # ------------------------------------------------------------------ generations
    def generation(self, plan, tree_pool, want_fitness=True):
        plan = np.ascontiguousarray(plan)
        assert plan.dtype == PLAN_DTYPE and len(plan) == self.P
        pool = np.ascontiguousarray(tree_pool, np.int32)
        if len(pool) == 0:
            pool = np.zeros(1, np.int32)
        if not want_fitness:
            self._ck(self.L.gsgp_generation(self.h, plan.ctypes.data, capi.iptr(pool), len(pool), None, None, None))
            return None
        fit, fit_test, best = np.empty(self.P), np.empty(self.P), C.c_int32()
        self._ck(self.L.gsgp_generation(self.h, plan.ctypes.data, capi.iptr(pool), len(pool),
                                        capi.dptr(fit), capi.dptr(fit_test), C.byref(best)))
        return fit, fit_test, best.value

    def generation_compiled(self, plan, tree_pool, programs, prog_off, prog_desc):
        """gsgp_generation_compiled: the trees arrive already compiled (HostDriver.compile_plan / Planner.programs)"""
        plan = np.ascontiguousarray(plan)
        assert plan.dtype == PLAN_DTYPE and len(plan) == self.P
        pool = np.ascontiguousarray(tree_pool, np.int32)
        if len(pool) == 0:
            pool = np.zeros(1, np.int32)
        programs = np.ascontiguousarray(programs, np.uint32)
        prog_off, prog_desc = np.ascontiguousarray(prog_off, np.int32), np.ascontiguousarray(prog_desc, np.int32)
        assert len(prog_off) == 2 * self.P and len(prog_desc) == 2 * self.P and len(programs) % 2 == 0
        fit, fit_test, best = np.empty(self.P), np.empty(self.P), C.c_int32()
        self._ck(self.L.gsgp_generation_compiled(self.h, plan.ctypes.data, capi.iptr(pool), len(pool),
                                                 programs.ctypes.data if len(programs) else pool.ctypes.data, len(programs) // 2,
                                                 capi.iptr(prog_off), capi.iptr(prog_desc),
                                                 capi.dptr(fit), capi.dptr(fit_test), C.byref(best)))
        return fit, fit_test, best.value

    def run_generations(self, n):
        self._ck(self.L.gsgp_run_generations(self.h, n))

    def sync(self):
        self._ck(self.L.gsgp_sync(self.h))

    # ------------------------------------------------------------------ outputs
    def get_fitness(self):
        fit, fit_test, best = np.empty(self.P), np.empty(self.P), C.c_int32()
        self._ck(self.L.gsgp_get_fitness(self.h, capi.dptr(fit), capi.dptr(fit_test), C.byref(best)))
        return fit, fit_test, best.value

    def get_semantic(self, ind):
        out = np.empty(self.n_local)
        self._ck(self.L.gsgp_get_semantic(self.h, ind, capi.dptr(out)))
        return out

    def get_semantics(self):
        return np.stack([self.get_semantic(i) for i in range(self.P)])

    def get_tree_semantic(self, ind, which):
        out = np.empty(self.n_local)
        self._ck(self.L.gsgp_get_tree_semantic(self.h, ind, which, capi.dptr(out)))
        return out

    def get_plan(self):
        plan = np.zeros(self.P, PLAN_DTYPE)
        self._ck(self.L.gsgp_get_plan(self.h, plan.ctypes.data))
        return plan

    def get_tree(self, ind, which):
        cap = 4 * (1 << max(self.cfg.max_depth_creation, 1)) + 8
        out, n = np.zeros(cap, np.int32), C.c_int32()
        self._ck(self.L.gsgp_get_tree(self.h, ind, which, capi.iptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def get_history(self, first, n):
        bf, bt, bi = np.empty(n), np.empty(n), np.empty(n, np.int32)
        self._ck(self.L.gsgp_get_history(self.h, first, n, capi.dptr(bf), capi.dptr(bt), capi.iptr(bi)))
        return bf, bt, bi

    def get_best_semantic_of(self, generation):
        out = np.empty(self.n_local)
        self._ck(self.L.gsgp_get_best_semantic_of(self.h, generation, capi.dptr(out)))
        return out

    def get_plan_of(self, generation):
        plan = np.zeros(self.P, PLAN_DTYPE)
        self._ck(self.L.gsgp_get_plan_of(self.h, generation, plan.ctypes.data))
        return plan

    @property
    def generation_index(self):
        return self.L.gsgp_generation_index(self.h)

    # ------------------------------------------------------------------ primitives
    def eval_trees(self, trees):
        pool, off, ln = pack_trees(trees)
        out = np.empty((len(trees), self.n_local))
        self._ck(self.L.gsgp_eval_trees(self.h, len(trees), capi.iptr(pool), len(pool), capi.iptr(off),
                                        capi.iptr(ln), capi.dptr(out)))
        return out

    def set_fitness(self, fit, fit_test=None):
        fit = np.ascontiguousarray(fit, np.float64)
        ft = None if fit_test is None else np.ascontiguousarray(fit_test, np.float64)
        self._ck(self.L.gsgp_set_fitness(self.h, capi.dptr(fit), capi.dptr(ft)))

    def draw_plan(self, generation):
        plan = np.zeros(self.P, PLAN_DTYPE)
        self._ck(self.L.gsgp_draw_plan(self.h, generation, plan.ctypes.data))
        return plan

    # ------------------------------------------------------------------ measurement
    @property
    def stream(self) -> int:
        return int(self.L.gsgp_stream(self.h) or 0)

    def profile_enable(self, on=True):
        self._ck(self.L.gsgp_profile_enable(self.h, int(on)))

    def profile_read(self, kernel):
        n, ms, b = C.c_int64(), C.c_double(), C.c_double()
        self._ck(self.L.gsgp_profile_read(self.h, kernel, C.byref(n), C.byref(ms), C.byref(b)))
        return n.value, ms.value, b.value

    @property
    def launch_count(self):
        return self.L.gsgp_launch_count(self.h)
