"""Binding of include/gsgp_host.h: the host-side driver pieces (RNG, constants, initial trees,
per-generation plan) used by the Python harness and mirrored by the C++ CLI driver."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import PLAN_DTYPE


class HostParams(C.Structure):
    _fields_ = [
        ("population_size", C.c_int32), ("nvar", C.c_int32), ("num_constants", C.c_int32),
        ("max_depth_creation", C.c_int32), ("tournament_size", C.c_int32),
        ("zero_depth", C.c_int32), ("minimization_problem", C.c_int32), ("init_type", C.c_int32),
        ("p_crossover", C.c_double), ("p_mutation", C.c_double),
        ("min_random_constant", C.c_double), ("max_random_constant", C.c_double),
    ]


HOST_SIGNATURES = {
    "gsgph_rng_new": (C.c_void_p, [C.c_int64]),
    "gsgph_rng_free": (None, [C.c_void_p]),
    "gsgph_rng_float64": (C.c_double, [C.c_void_p]),
    "gsgph_rng_intn": (C.c_int32, [C.c_void_p, C.c_int32]),
    "gsgph_create_T_F": (C.c_int, [C.c_void_p, C.POINTER(HostParams), C.POINTER(C.c_double)]),
    "gsgph_initialize_population": (C.c_int, [C.c_void_p, C.POINTER(HostParams), C.c_int32, C.POINTER(C.c_int32),
                                              C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gsgph_build_plan": (C.c_int, [C.c_void_p, C.POINTER(HostParams), C.POINTER(C.c_double), C.c_int32, C.c_void_p,
                                   C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "gsgph_best_individual": (C.c_int32, [C.POINTER(HostParams), C.POINTER(C.c_double)]),
    "gsgph_planner_new": (C.c_void_p, [C.c_void_p, C.POINTER(HostParams)]),
    "gsgph_planner_free": (None, [C.c_void_p]),
    "gsgph_planner_prefetch": (C.c_int, [C.c_void_p, C.c_int32]),
    "gsgph_planner_plan": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_void_p),
                                     C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]),
    "gsgph_planner_enable_programs": (C.c_int, [C.c_void_p, C.c_int32]),
    "gsgph_planner_programs": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int32),
                                         C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32))]),
    "gsgph_compile_plan": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32),
                                     C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gsgph_replay_generations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gsgph_planner_stats": (None, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gsgph_build_effects_plan": (C.c_int, [C.c_void_p, C.POINTER(HostParams), C.c_int32, C.c_void_p,
                                           C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "gsgph_format_float": (C.c_int, [C.c_double, C.c_char_p, C.c_int32]),
    "gsgph_format_semantic": (C.c_int, [C.POINTER(C.c_double), C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.c_int32]),
    "gsgph_read_dataset": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.POINTER(C.c_double))]),
    "gsgph_read_semantic": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_double)]),
    "gsgph_free": (None, [C.c_void_p]),
    "gsgph_parse_floats": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64), C.c_int32]),
    "gsgph_compile_tree": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32), C.c_int32,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gsgph_writer_open": (C.c_void_p, [C.c_char_p, C.c_int32]),
    "gsgph_writer_semantic": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int64]),
    "gsgph_writer_float": (C.c_int, [C.c_void_p, C.c_double]),
    "gsgph_writer_close": (C.c_int, [C.c_void_p]),
    "gsgph_gzip_member": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "gsgph_format_duration": (C.c_int, [C.c_int64, C.c_char_p, C.c_int32]),
    "gsgph_channel_create": (C.c_void_p, [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gsgph_channel_attach": (C.c_void_p, [C.c_char_p, C.c_int32, C.c_int32]),
    "gsgph_channel_close": (None, [C.c_void_p]),
    "gsgph_channel_publish": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_int32,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]),
    "gsgph_channel_acquire": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_int32)),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                        C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)), C.c_int32]),
    "gsgph_channel_release": (C.c_int, [C.c_void_p, C.c_uint64]),
    "gsgph_replay_generations_shared": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double), C.c_int32]),
    "gsgph_contrib_new": (C.c_void_p, [C.c_int32, C.c_int32]),
    "gsgph_contrib_free": (None, [C.c_void_p]),
    "gsgph_contrib_num_models": (C.c_int32, [C.c_void_p]),
    "gsgph_contrib_limbs": (C.c_int32, [C.c_void_p]),
    "gsgph_contrib_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "gsgph_contrib_format": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "gsgph_proto_new": (C.c_void_p, []),
    "gsgph_proto_free": (None, [C.c_void_p]),
    "gsgph_proto_begin_population": (C.c_int, [C.c_void_p, C.c_int32]),
    "gsgph_proto_add_individual": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double,
                                             C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64,
                                             C.c_char_p, C.c_int64, C.c_int32, C.POINTER(C.POINTER(C.c_int32)),
                                             C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double))]),
    "gsgph_proto_end_population": (C.c_int, [C.c_void_p]),
    "gsgph_proto_size": (C.c_int64, [C.c_void_p]),
    "gsgph_proto_bytes": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "gsgph_proto_write": (C.c_int, [C.c_void_p, C.c_char_p]),
}

_bound = False


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        for name, (res, args) in HOST_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def format_float(v: float) -> str:
    buf = C.create_string_buffer(64)
    _lib().gsgph_format_float(float(v), buf, 64)
    return buf.value.decode()


def format_semantic(sem) -> str:
    """Semantic.String() (main_cpu.go:123-126): "3.14,2.71,1.41"."""
    sem = np.ascontiguousarray(sem, np.float64)
    cap = 25 * max(1, len(sem))
    buf = C.create_string_buffer(cap)
    n = C.c_int64()
    rc = _lib().gsgph_format_semantic(capi.dptr(sem), len(sem), buf, cap, C.byref(n), 0)
    assert rc == 0
    return buf.raw[: n.value].decode()


def read_dataset(train_path, test_path):
    """read_input_data (main_cpu.go:381-421) -> (X (N,V), y (N,), nrow); raises like the reference panics."""
    nv, nr, nt = C.c_int32(), C.c_int32(), C.c_int32()
    ptr = C.POINTER(C.c_double)()
    rc = _lib().gsgph_read_dataset(str(train_path).encode(), str(test_path).encode(), C.byref(nv), C.byref(nr), C.byref(nt), C.byref(ptr))
    if rc:
        raise RuntimeError({1: "cannot open dataset file", 3: "not enough values in dataset file",
                            4: "Train and Test datasets must have the same number of variables"}.get(rc, "cannot parse dataset file"))
    n, v = nr.value + nt.value, nv.value
    m = np.ctypeslib.as_array(ptr, shape=(n, v + 1)).copy()
    _lib().gsgph_free(ptr)
    return m[:, :v].copy(), m[:, v].copy(), nr.value


def read_semantic(path, n):
    """read_sem (main_cpu.go:710-733): n values, one per line."""
    out = np.empty(n)
    rc = _lib().gsgph_read_semantic(str(path).encode(), n, capi.dptr(out))
    if rc:
        raise RuntimeError({1: "cannot open semantic file", 2: "Cannot parse semantic value",
                            3: "Not enough values when reading semantic file"}[rc])
    return out


def compile_tree(tree, n_symbols):
    """Reference tree array -> (code[n], ab[n], spill_levels): the program the interpreter kernel runs."""
    t = np.ascontiguousarray(tree, np.int32)
    cap = max(1, len(t))
    out = np.zeros(2 * cap, np.uint32)
    n, need = C.c_int32(), C.c_int32()
    rc = _lib().gsgph_compile_tree(capi.iptr(t), len(t), n_symbols, out.ctypes.data_as(C.POINTER(C.c_uint32)), cap,
                                   C.byref(n), C.byref(need))
    if rc:
        raise ValueError(f"malformed tree (gsgph_compile_tree rc={rc})")
    return out[0:2 * n.value:2].copy(), out[1:2 * n.value:2].copy(), need.value


def format_duration(ns: int) -> str:
    """time.Duration.String(): what fmt.Fprintln(executiontime, time.Since(start)) prints (main_cpu.go:1413,1505)"""
    buf = C.create_string_buffer(40)
    n = _lib().gsgph_format_duration(int(ns), buf, 40)
    return buf.raw[:n].decode()


def gzip_member(data: bytes) -> bytes:
    """gsgph_gzip_member: data as one gzip member, the ".gz" writers' compressor"""
    cap = len(data) * 2 + 1024
    buf, n = C.create_string_buffer(cap), C.c_int64()
    rc = _lib().gsgph_gzip_member(data, len(data), buf, cap, C.byref(n))
    if rc:
        raise RuntimeError(f"gsgph_gzip_member failed ({rc})")
    return buf.raw[: n.value]


class PlanChannel:
    """gsgph_channel_*: one rank (the writer) plans, the other ranks of the node read plan + tree pool + programs from
    shared memory.  PlanChannel.create(...) on the writer, PlanChannel.attach(...) on each reader."""

    def __init__(self, h, P, writer):
        self.L, self.h, self.P, self.writer = _lib(), h, P, writer

    @classmethod
    def create(cls, name, population_size, cap_pool_ints, cap_instructions, n_readers):
        h = _lib().gsgph_channel_create(os.fsencode(name), population_size, cap_pool_ints, cap_instructions, n_readers)
        if not h:
            raise OSError(f"gsgph_channel_create failed: {name}")
        return cls(h, population_size, True)

    @classmethod
    def attach(cls, name, reader_index, population_size, timeout_ms=10000):
        h = _lib().gsgph_channel_attach(os.fsencode(name), reader_index, timeout_ms)
        if not h:
            raise OSError(f"gsgph_channel_attach failed: {name}")
        return cls(h, population_size, False)

    def close(self):
        if self.h:
            self.L.gsgph_channel_close(self.h)
            self.h = None

    def publish(self, seq, plan, pool, programs, prog_off, prog_desc, timeout_ms=10000):
        plan = np.ascontiguousarray(plan, PLAN_DTYPE)
        pool = np.ascontiguousarray(pool, np.int32)
        programs = np.ascontiguousarray(programs, np.uint32)
        prog_off, prog_desc = np.ascontiguousarray(prog_off, np.int32), np.ascontiguousarray(prog_desc, np.int32)
        rc = self.L.gsgph_channel_publish(self.h, seq, plan.ctypes.data, capi.iptr(pool), len(pool), programs.ctypes.data,
                                          len(programs) // 2, capi.iptr(prog_off), capi.iptr(prog_desc), timeout_ms)
        if rc:
            raise RuntimeError(f"gsgph_channel_publish failed ({rc})")

    def acquire(self, seq, timeout_ms=10000):
        """-> (plan, pool, programs, prog_off, prog_desc): copies of what the writer published as `seq`"""
        plan_p, prog_p = C.c_void_p(), C.c_void_p()
        pool_p, off_p, desc_p = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        n_pool, n_ins = C.c_int32(), C.c_int32()
        rc = self.L.gsgph_channel_acquire(self.h, seq, C.byref(plan_p), C.byref(pool_p), C.byref(n_pool), C.byref(prog_p),
                                          C.byref(n_ins), C.byref(off_p), C.byref(desc_p), timeout_ms)
        if rc:
            raise RuntimeError(f"gsgph_channel_acquire failed ({rc})")
        view = lambda addr, ctype, n, dt: np.ctypeslib.as_array((ctype * max(n, 1)).from_address(addr)).view(dt)[:n].copy()
        plan = np.frombuffer((C.c_char * (40 * self.P)).from_address(plan_p.value), PLAN_DTYPE).copy()
        return (plan, view(C.addressof(pool_p.contents), C.c_int32, n_pool.value, np.int32),
                view(prog_p.value, C.c_uint32, 2 * n_ins.value, np.uint32),
                view(C.addressof(off_p.contents), C.c_int32, 2 * self.P, np.int32),
                view(C.addressof(desc_p.contents), C.c_int32, 2 * self.P, np.int32))

    def release(self, seq):
        self.L.gsgph_channel_release(self.h, seq)

    def replay_generations(self, planner, engine, first_seq, n, fit, fit_test, index_best, timeout_ms=60000):
        """gsgph_replay_generations_shared: the writer passes its Planner (programs enabled), readers pass None.
        Returns (fit, fit_test, best, best_fit_history, best_fit_test_history)."""
        fit, fit_test = np.array(fit, np.float64), np.array(fit_test, np.float64)
        best = C.c_int32(int(index_best))
        h0, h1 = np.empty(max(n, 1)), np.empty(max(n, 1))
        rc = self.L.gsgph_replay_generations_shared(planner.h if planner is not None else None, self.h, engine.h, first_seq, n,
                                                    capi.dptr(fit), capi.dptr(fit_test), C.byref(best), capi.dptr(h0), capi.dptr(h1),
                                                    timeout_ms)
        if rc == 4:
            engine._ck(1)
        if rc:
            raise RuntimeError(f"gsgph_replay_generations_shared failed ({rc})")
        return fit, fit_test, best.value, h0[:n], h1[:n]


class Contributions:
    """contrib / contrib_new of main_cpu.go (:128-149,215-216): per individual, how often each seeded model (slot i + 1)
    and GP itself (slot 0) appears among its ancestors.  apply(plan) is one generation (:857,885,909,1196)."""

    def __init__(self, population_size, n_seeds=0):
        self.L = _lib()
        self.h = self.L.gsgph_contrib_new(population_size, n_seeds)
        if not self.h:
            raise ValueError("gsgph_contrib_new: bad sizes")
        self.P = population_size

    def close(self):
        if self.h:
            self.L.gsgph_contrib_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_models(self):
        return self.L.gsgph_contrib_num_models(self.h)

    @property
    def limbs(self):
        return self.L.gsgph_contrib_limbs(self.h)

    def apply(self, plan, n_threads=0):
        plan = np.ascontiguousarray(plan, PLAN_DTYPE)
        assert len(plan) == self.P
        rc = self.L.gsgph_contrib_apply(self.h, plan.ctypes.data, n_threads)
        if rc:
            raise ValueError(f"gsgph_contrib_apply failed ({rc})")

    def format(self, ind) -> str:
        """Contribution.String() (:141-146), e.g. "3,1,4" """
        n = C.c_int64()
        self.L.gsgph_contrib_format(self.h, ind, None, 0, C.byref(n))
        buf = C.create_string_buffer(max(n.value, 1))
        rc = self.L.gsgph_contrib_format(self.h, ind, buf, n.value, C.byref(n))
        if rc:
            raise ValueError(f"gsgph_contrib_format failed ({rc})")
        return buf.raw[: n.value].decode()


class ProtoDump:
    """-proto_dump (main_cpu.go:1416-1436,1441-1488,1510-1517): pb.Evolution, one pb.Population per generation, written
    in protobuf wire format by gsgph_proto_* (readable with the reference's evolution.proto).  Semantics are passed as
    one row per individual, train cases then test cases (the store's layout)."""

    def __init__(self, nrow, nrow_test):
        self.L, self.nrow, self.nrow_test = _lib(), int(nrow), int(nrow_test)
        self.h = self.L.gsgph_proto_new()
        self._rt = [None, None]          # rt1 / rt2 and their semantics are globals in the reference (:217-226)

    def close(self):
        if self.h:
            self.L.gsgph_proto_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin_population(self, generation):
        if self.L.gsgph_proto_begin_population(self.h, int(generation)):
            raise RuntimeError("gsgph_proto_begin_population: a population is already open")

    def end_population(self):
        if self.L.gsgph_proto_end_population(self.h):
            raise RuntimeError("gsgph_proto_end_population: no population is open")

    def add_individual(self, op, fit, fit_test, semantic, contrib=b"", trees=()):
        """trees: up to two (data, semantic) pairs -- the operator's random trees; (None, None) is the reference's
        not-yet-assigned rt (an empty RandomTree)"""
        sem = np.ascontiguousarray(semantic, np.float64)
        assert sem.shape == (self.nrow + self.nrow_test,)
        n = len(trees)
        keep, data = [], (C.POINTER(C.c_int32) * max(n, 1))()
        lens = (C.c_int32 * max(n, 1))()
        tr, te = (C.POINTER(C.c_double) * max(n, 1))(), (C.POINTER(C.c_double) * max(n, 1))()
        for t, (d, ts) in enumerate(trees):
            if d is None:
                continue                                         # NULL pointers, length 0
            d = np.ascontiguousarray(d, np.int32)
            ts = np.ascontiguousarray(ts, np.float64)
            assert ts.shape == sem.shape
            keep += [d, ts]
            data[t], lens[t] = capi.iptr(d), len(d)
            tr[t] = capi.dptr(ts)
            te[t] = C.cast(C.c_void_p(ts.ctypes.data + 8 * self.nrow), C.POINTER(C.c_double))
        test = C.cast(C.c_void_p(sem.ctypes.data + 8 * self.nrow), C.POINTER(C.c_double))
        rc = self.L.gsgph_proto_add_individual(self.h, int(op), float(fit), float(fit_test), capi.dptr(sem), self.nrow, test,
                                               self.nrow_test, bytes(contrib), len(contrib), n, data, lens, tr, te)
        if rc:
            raise RuntimeError("gsgph_proto_add_individual: bad arguments or no open population")

    def add_initial_population(self, fit, fit_test, semantics, contribs=None):
        """generation 0 (:1418-1436): every individual is Individual_INIT, no random trees"""
        self.begin_population(0)
        for k in range(len(fit)):
            self.add_individual(capi.OP_INIT, fit[k], fit_test[k], semantics[k], contribs[k] if contribs else b"")
        self.end_population()

    def add_generation(self, generation, plan, pool, fit_new, fit_test_new, semantics_new, tree_semantic, contribs=None, tree=None):
        """one generation's loop (:1441-1488) from the plan that produced it.  tree_semantic(k, which) -> the semantic of
        individual k's random tree (Engine.get_tree_semantic with GSGP_FLAG_KEEP_TREE_SEMANTICS).  The elite creates no
        tree, so -- as in the reference, where rt1 / rt2 are globals -- its entry repeats the previous individual's.
        contribs: Contribution.String() bytes per individual; the reference passes contrib[k] of the generation being
        replaced (:1479, before update_tables swaps the tables).  tree(k, which) -> the tree array, instead of `pool`
        (device mode keeps the trees on the device: Engine.get_tree)."""
        self.begin_population(generation)
        for k in range(len(plan)):
            e = plan[k]
            op = int(e["op"])
            if op in (capi.OP_XO, capi.OP_MUT) and not e["elite"]:
                arr = tree if tree is not None else (lambda k_, w_: pool[e[f"tree{w_}_off"]: e[f"tree{w_}_off"] + e[f"tree{w_}_len"]].copy())
                self._rt[0] = (arr(k, 0), tree_semantic(k, 0))
                if op == capi.OP_MUT:
                    self._rt[1] = (arr(k, 1), tree_semantic(k, 1))
            trees = [r if r is not None else (None, None) for r in self._rt[: {capi.OP_XO: 1, capi.OP_MUT: 2}.get(op, 0)]]
            self.add_individual(op, fit_new[k], fit_test_new[k], semantics_new[k], contribs[k] if contribs else b"", trees)
        self.end_population()

    def tobytes(self) -> bytes:
        n = self.L.gsgph_proto_size(self.h)
        buf = C.create_string_buffer(max(n, 1))
        if self.L.gsgph_proto_bytes(self.h, buf, n):
            raise RuntimeError("gsgph_proto_bytes: a population is still open")
        return buf.raw[:n]

    def write(self, path):
        rc = self.L.gsgph_proto_write(self.h, os.fsencode(path))
        if rc:
            raise OSError(f"gsgph_proto_write failed ({rc}): {path}")


class LineWriter:
    """The reference's per-generation output files (main_cpu.go:1227-1241,1494-1500): fitness lines and
    best-individual semantic lines, gzip when the path ends in ".gz".  Writes are asynchronous: the values are
    copied, a background thread formats and compresses."""

    def __init__(self, path, n_threads=0):
        self.L = _lib()
        self.h = self.L.gsgph_writer_open(str(path).encode(), n_threads)
        if not self.h:
            raise OSError(f"cannot create {path}")

    def write_semantic(self, sem):
        sem = np.ascontiguousarray(sem, np.float64)
        if self.L.gsgph_writer_semantic(self.h, capi.dptr(sem), len(sem)):
            raise OSError("semantic line could not be queued")

    def write_float(self, v):
        if self.L.gsgph_writer_float(self.h, float(v)):
            raise OSError("line could not be queued")

    def close(self):
        if self.h:
            rc = self.L.gsgph_writer_close(self.h)
            self.h = None
            if rc:
                raise OSError("output file could not be written completely")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Planner:
    """gsgph_planner_*: build_plan's result with the fitness-independent draws made on a worker thread while the
    device computes the previous generation.  plan(fit, best) -> (plan, pool) views valid until the next plan();
    prefetch(best) starts the following generation's draws.  With programs=True the trees are also compiled ahead
    of time; programs() returns what Engine.generation_compiled takes."""

    def __init__(self, driver: "HostDriver", programs=False, n_threads=0):
        self.L = driver.L
        self.driver = driver                                     # keeps the rng alive
        self.h = self.L.gsgph_planner_new(driver.rng, C.byref(driver.p))
        if not self.h:
            raise ValueError("gsgph_planner_new: bad parameters")
        self.P = driver.P
        self._views = {}                                         # the planner rotates three buffers: views by address
        if programs and self.L.gsgph_planner_enable_programs(self.h, n_threads):
            raise RuntimeError("gsgph_planner_enable_programs failed")

    def _view(self, addr, ctype, count, dtype):
        key = (addr, count)
        v = self._views.get(key)
        if v is None:
            if len(self._views) > 64:
                self._views.clear()
            v = np.frombuffer((ctype * count).from_address(addr), dtype=dtype) if count else np.zeros(0, dtype)
            self._views[key] = v
        return v

    def plan(self, fit, index_best):
        fit = np.ascontiguousarray(fit, np.float64)
        plan_p, pool_p, n = C.c_void_p(), C.POINTER(C.c_int32)(), C.c_int32()
        rc = self.L.gsgph_planner_plan(self.h, capi.dptr(fit), int(index_best), C.byref(plan_p), C.byref(pool_p), C.byref(n))
        if rc:
            raise RuntimeError(f"gsgph_planner_plan failed ({rc})")
        plan = self._view(plan_p.value, C.c_char, PLAN_DTYPE.itemsize * self.P, PLAN_DTYPE)
        pool = self._view(C.addressof(pool_p.contents), C.c_int32, max(n.value, 1), np.int32)
        return plan, pool

    def programs(self):
        """(programs uint32[2*n], prog_off int32[2P], prog_desc int32[2P]) of the plan returned last"""
        pr, n = C.POINTER(C.c_uint32)(), C.c_int32()
        off, desc = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        if self.L.gsgph_planner_programs(self.h, C.byref(pr), C.byref(n), C.byref(off), C.byref(desc)):
            raise RuntimeError("planner was created without programs=True")
        return (self._view(C.addressof(pr.contents), C.c_uint32, 2 * max(n.value, 1), np.uint32)[: 2 * n.value],
                self._view(C.addressof(off.contents), C.c_int32, 2 * self.P, np.int32),
                self._view(C.addressof(desc.contents), C.c_int32, 2 * self.P, np.int32))

    def replay_generations(self, engine, n, fit, fit_test, index_best):
        """gsgph_replay_generations: n generations of the host-replay loop in the compiled host (no Python in between).
        Returns (fit, fit_test, best, best_fit_history, best_fit_test_history)."""
        fit, fit_test = np.array(fit, np.float64), np.array(fit_test, np.float64)
        best = C.c_int32(int(index_best))
        h0, h1 = np.empty(max(n, 1)), np.empty(max(n, 1))
        rc = self.L.gsgph_replay_generations(self.h, engine.h, n, capi.dptr(fit), capi.dptr(fit_test), C.byref(best),
                                             capi.dptr(h0), capi.dptr(h1))
        if rc == 4:
            engine._ck(1)
        if rc:
            raise RuntimeError(f"gsgph_replay_generations failed ({rc})")
        return fit, fit_test, best.value, h0[:n], h1[:n]

    def prefetch(self, index_best_guess):
        rc = self.L.gsgph_planner_prefetch(self.h, int(index_best_guess))
        if rc:
            raise RuntimeError(f"gsgph_planner_prefetch failed ({rc})")

    def stats(self):
        h, m, r = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.gsgph_planner_stats(self.h, C.byref(h), C.byref(m), C.byref(r))
        return {"hits": h.value, "misses": m.value, "redrawn_individuals": r.value}

    def close(self):
        if self.h:
            self.L.gsgph_planner_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostDriver:
    """The RNG-order-defining half of main_cpu.go's main(): seeds the RNG, draws the constants,
    builds the initial population and, each generation, the plan the device applies."""

    def __init__(self, *, population_size, nvar, num_constants=0, max_depth_creation=6, tournament_size=4,
                 zero_depth=False, minimization_problem=True, init_type=2, p_crossover=0.6, p_mutation=0.3,
                 min_random_constant=-100.0, max_random_constant=100.0, seed=1):
        self.L = _lib()
        self.p = HostParams(population_size, nvar, num_constants, max_depth_creation, tournament_size,
                            int(zero_depth), int(minimization_problem), init_type, p_crossover, p_mutation,
                            min_random_constant, max_random_constant)
        self.rng = self.L.gsgph_rng_new(seed)                    # rand.Seed main_cpu.go:1367
        self.P = population_size
        d = max(max_depth_creation, 1)
        self._tree_cap = 4 * (1 << d)
        self._plan = np.zeros(self.P, PLAN_DTYPE)
        self._pool = np.zeros(2 * self.P * self._tree_cap, np.int32)

    def close(self):
        for pl in getattr(self, "_planners", []):               # a planner's worker draws from this rng
            pl.close()
        if self.rng:
            self.L.gsgph_rng_free(self.rng)
            self.rng = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_T_F(self) -> np.ndarray:
        sv = np.zeros(4 + self.p.nvar + self.p.num_constants)
        self.L.gsgph_create_T_F(self.rng, C.byref(self.p), capi.dptr(sv))
        return sv

    def initialize_population(self, n_seeds=0):
        n = self.P - n_seeds
        cap = max(1, n) * (self._tree_cap + 4)
        pool = np.zeros(cap, np.int32)
        off, ln, used = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), C.c_int32()
        rc = self.L.gsgph_initialize_population(self.rng, C.byref(self.p), n_seeds, capi.iptr(pool), cap,
                                                capi.iptr(off), capi.iptr(ln), C.byref(used))
        if rc:
            raise RuntimeError(f"gsgph_initialize_population failed ({rc})")
        return [pool[off[i]: off[i] + ln[i]].copy() for i in range(n)]

    def build_plan(self, fit: np.ndarray, index_best: int):
        """Returns (plan, pool) views valid until the next call."""
        fit = np.ascontiguousarray(fit, np.float64)
        used = C.c_int32()
        rc = self.L.gsgph_build_plan(self.rng, C.byref(self.p), capi.dptr(fit), index_best, self._plan.ctypes.data,
                                     capi.iptr(self._pool), len(self._pool), C.byref(used))
        if rc:
            raise RuntimeError("gsgph_build_plan: tree pool overflow")
        return self._plan, self._pool[: max(used.value, 1)]

    def planner(self, programs=False, n_threads=0) -> Planner:
        pl = Planner(self, programs, n_threads)
        self.__dict__.setdefault("_planners", []).append(pl)
        return pl

    def compile_plan(self, plan, pool):
        """gsgph_compile_plan: (programs, prog_off, prog_desc) for Engine.generation_compiled"""
        pool = np.ascontiguousarray(pool, np.int32)
        cap = len(pool) // 4 + 2 * self.P + 1
        prog, n = np.zeros(2 * cap, np.uint32), C.c_int32()
        off, desc = np.zeros(2 * self.P, np.int32), np.zeros(2 * self.P, np.int32)
        rc = self.L.gsgph_compile_plan(plan.ctypes.data, self.P, capi.iptr(pool), len(pool), 4 + self.p.nvar + self.p.num_constants,
                                       prog.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(n), capi.iptr(off), capi.iptr(desc))
        if rc:
            raise ValueError(f"gsgph_compile_plan failed ({rc})")
        return prog[: 2 * n.value], off, desc

    def build_effects_plan(self, test_crossover: bool):
        """operators_effects.go's single step (XO-only with uniform parents, or self-copy + mutation) as a plan."""
        used = C.c_int32()
        rc = self.L.gsgph_build_effects_plan(self.rng, C.byref(self.p), int(test_crossover), self._plan.ctypes.data,
                                             capi.iptr(self._pool), len(self._pool), C.byref(used))
        if rc:
            raise RuntimeError("gsgph_build_effects_plan: tree pool overflow")
        return self._plan, self._pool[: max(used.value, 1)]

    def best_individual(self, fit) -> int:
        fit = np.ascontiguousarray(fit, np.float64)
        return self.L.gsgph_best_individual(C.byref(self.p), capi.dptr(fit))
