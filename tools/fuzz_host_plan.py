"""Differential fuzz of the two restatements of main_cpu.go's RNG-order-defining half: the product's host driver
(gsgph_*: serial plan builder and pipelined planner) against the oracle, under Go's rand.Seed stream, on random
configurations (population, depth, constants, tournament size, init type, operator rates, seeded individuals, seeds
incl. negative).  Runs for about 150 s.  Last run (end of round 2): 43 014 configurations, 0 differences.
    python tools/fuzz_host_plan.py"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as ob
from tests.util import synth_dataset
from go_gsgp_b200.host import HostDriver
X, y, nrow = synth_dataset(1, 120, 7, nrow=90)
t0 = time.time(); n = 0; bad = 0
rng = np.random.default_rng(12345)
while time.time() - t0 < 150:
    kw = dict(population_size=int(rng.integers(1, 80)), num_random_constants=int(rng.choice([0, 0, 1, 5])),
              max_depth_creation=int(rng.integers(1, 9)), tournament_size=int(rng.integers(1, 8)), zero_depth=int(rng.integers(0, 2)),
              minimization_problem=int(rng.integers(0, 2)), init_type=int(rng.integers(0, 3)), p_crossover=float(rng.choice([0, .3, .6, 1])))
    kw["p_mutation"] = float(min(1 - kw["p_crossover"], rng.choice([0, .3, .4, 1])))
    seed = int(rng.integers(-2**40, 2**40))
    n_seeds = int(rng.integers(0, kw["population_size"] + 1)) if rng.random() < 0.3 else 0
    try:
        o = ob.Oracle(ob.make_config(rng_kind=ob.RNG_ALFG, rng_seed=seed, **kw), X, y, nrow); o.create_T_F()
        h = HostDriver(nvar=7, seed=seed, population_size=kw["population_size"], num_constants=kw["num_random_constants"],
                       max_depth_creation=kw["max_depth_creation"], tournament_size=kw["tournament_size"], zero_depth=kw["zero_depth"],
                       minimization_problem=kw["minimization_problem"], init_type=kw["init_type"], p_crossover=kw["p_crossover"], p_mutation=kw["p_mutation"])
        assert np.array_equal(h.create_T_F(), o.symbol_values())
        for i in range(n_seeds): o.seed_semantic(i, y + 0.01 * i)
        o.initialize_population(n_seeds); trees = h.initialize_population(n_seeds)
        assert len(trees) == o.P - n_seeds
        for i, t in enumerate(trees): assert np.array_equal(t, o.initial_tree(n_seeds + i))
        o.evaluate()
        pl = h.planner() if rng.random() < 0.5 else None
        for g in range(5):
            fit, best = o.fit(), o.index_best
            assert h.best_individual(fit) == best
            if pl is None: plan, pool = h.build_plan(fit, best)
            else:
                plan, pool = pl.plan(fit, best); pl.prefetch(best)
            o.generation(); want = o.last_plan()
            assert plan.tobytes() == want.tobytes(), (g, "plan")
            assert np.array_equal(pool[:len(o.last_tree_pool())], o.last_tree_pool()), (g, "pool")
        h.close(); o.close()
    except Exception as ex:
        bad += 1; print("FAIL", kw, seed, n_seeds, repr(ex)[:300])
        if bad > 5: break
    n += 1
print("configs", n, "bad", bad)
