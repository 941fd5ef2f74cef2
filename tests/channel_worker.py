"""Child process of tests/test_plan_channel.py (not a test module).
    reader  <name> <index> <P> <n_seq> <slow>      attach, acquire/hash/release n_seq plans, print the hashes
    replay  <stub.so> <role> <name> <P> <gens>     gsgph_replay_generations_shared over the stub device (LD_PRELOAD'ed):
                                                   role = writer | reader<i> | solo (plain gsgph_replay_generations)"""
import ctypes as C
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                            # noqa: E402
from go_gsgp_b200 import capi                                 # noqa: E402
from go_gsgp_b200.host import HostDriver, PlanChannel, _lib   # noqa: E402


def digest(parts):
    h = hashlib.sha256()
    for a in parts:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def reader(name, index, P, n_seq, slow):
    ch = PlanChannel.attach(name, index, P, timeout_ms=20000)
    out = []
    for seq in range(1, n_seq + 1):
        parts = ch.acquire(seq, timeout_ms=20000)
        out.append(digest(parts))
        if slow and seq % 5 == 0:
            time.sleep(0.02)                                  # the writer has to wait for this reader's release
        ch.release(seq)
    ch.close()
    print(json.dumps(out))


def replay(stub_path, role, name, P, gens):
    stub = C.CDLL(stub_path)                                  # gsgp_create etc. of the stub (also LD_PRELOAD'ed, so the
    L = _lib()                                                # library's own call to gsgp_generation_compiled lands there)
    cfg = capi.GsgpConfig()
    cfg.population_size, cfg.nvar, cfg.nrow, cfg.nrow_test = P, 5, 40, 10
    ctx = C.c_void_p()
    stub.gsgp_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    assert stub.gsgp_create(C.byref(cfg), C.byref(ctx)) == 0
    fit, fit_test, best = np.zeros(P), np.zeros(P), C.c_int32()
    stub.gsgp_fitness_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    stub.gsgp_fitness_all(ctx, fit.ctypes.data, fit_test.ctypes.data, C.byref(best))
    h0, h1 = np.zeros(gens), np.zeros(gens)
    host = pl = ch = None
    if role in ("writer", "solo"):
        host = HostDriver(population_size=P, nvar=5, seed=7)
        host.create_T_F(); host.initialize_population(0)
        pl = host.planner(programs=True)
    if role == "solo":
        rc = L.gsgph_replay_generations(pl.h, ctx, gens, capi.dptr(fit), capi.dptr(fit_test), C.byref(best), capi.dptr(h0), capi.dptr(h1))
    else:
        if role == "writer":
            ch = PlanChannel.create(name, P, 2 * P * 260, 2 * P * 70, int(os.environ["N_READERS"]))
        else:
            ch = PlanChannel.attach(name, int(role[len("reader"):]), P, timeout_ms=20000)
        rc = 0
        for first, n in ((1, gens // 2), (1 + gens // 2, gens - gens // 2)):       # two calls: first_seq continues
            rc = rc or L.gsgph_replay_generations_shared(pl.h if pl else None, ch.h, ctx, first, n, capi.dptr(fit), capi.dptr(fit_test),
                                                         C.byref(best), capi.dptr(h0[first - 1:]), capi.dptr(h1[first - 1:]), 20000)
    stub.stub_plan_hash.restype = C.c_uint64
    stub.stub_plan_hash.argtypes = [C.c_void_p]
    print(json.dumps({"rc": rc, "hash": stub.stub_plan_hash(ctx), "best": best.value, "hist": h0.tolist(), "fit": fit.tolist()}))
    if role == "writer":
        time.sleep(0.2)                                       # let the readers finish before the segment is unlinked
    if ch:
        ch.close()
    if host:
        host.close()


if __name__ == "__main__":
    if sys.argv[1] == "reader":
        reader(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6] == "1")
    else:
        replay(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]))
