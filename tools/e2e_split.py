"""Where a host-replay generation's wall time goes at config 2: plan building (serial gsgph_build_plan vs the
pipelined gsgph_planner_*), the gsgp_generation() call, and gen_kernel inside it.  GSGP_TRACE=1 adds the call's
own breakdown (validate / stage+upload+launch / finalize launches / wait+readback) on stderr."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_gsgp_b200 import Engine, capi
from go_gsgp_b200.host import HostDriver
from bench import synth

P, V, N = int(os.environ.get("POP", 1000)), 20, int(os.environ.get("CASES", 100000))
X, y = synth(2, N, V); nrow = int(0.8 * N)
for mode in ("serial", "pipelined", "programs"):
    host = HostDriver(population_size=P, nvar=V, seed=1); sym = host.create_T_F(); trees = host.initialize_population(0)
    e = Engine(population_size=P, nvar=V, nrow=nrow, nrow_test=N - nrow, error_measure=capi.RMSE)
    e.set_dataset(X, y); e.set_constants(sym); e.eval_initial(0, trees); fit, ft, best = e.fitness_all()
    pl = host.planner(programs=mode == "programs") if mode != "serial" else None
    tb = tg = 0.0
    for i in range(110):
        t0 = time.perf_counter()
        if pl is None:
            plan, pool = host.build_plan(fit, best)
        else:
            plan, pool = pl.plan(fit, best); pl.prefetch(best)
        t1 = time.perf_counter()
        fit, ft, best = (e.generation_compiled(plan, pool, *pl.programs()) if mode == "programs" else e.generation(plan, pool)); t2 = time.perf_counter()
        if i >= 10: tb += t1 - t0; tg += t2 - t1
    print('%-9s plan %.3f ms  generation() %.3f ms  total %.3f ms/gen  %s' % (mode, tb / 100 * 1e3, tg / 100 * 1e3, (tb + tg) / 100 * 1e3,
          pl.stats() if pl else ''), flush=True)
    e.profile_enable(True)
    for i in range(20):
        plan, pool = (host.build_plan(fit, best) if pl is None else pl.plan(fit, best))
        if pl: pl.prefetch(best)
        fit, ft, best = (e.generation_compiled(plan, pool, *pl.programs()) if mode == "programs" else e.generation(plan, pool))
    n, ms, _ = e.profile_read(capi.K_GSOP); print('          gen kernel ms', ms / n, flush=True)
    e.close(); host.close()
