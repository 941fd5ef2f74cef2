"""The C++ host-side driver pieces (include/gsgp_host.h) against the oracle's restatement of the
same reference code: two independent implementations of main_cpu.go's RNG-order-defining logic must
produce identical constants, initial trees and generation plans from the same seed."""
import numpy as np
import pytest

from oracle import binding as ob
from tests.util import synth_dataset


@pytest.fixture(scope="module", autouse=True)
def _built():
    from go_gsgp_b200 import build
    build.build()


@pytest.mark.parametrize("kw", [
    dict(population_size=100),
    dict(population_size=57, num_random_constants=6, max_depth_creation=8, tournament_size=3, zero_depth=1),
    dict(population_size=30, init_type=0, p_crossover=0.2, p_mutation=0.7, minimization_problem=0),
    dict(population_size=30, init_type=1, max_depth_creation=4),
])
def test_host_plan_equals_oracle_plan(kw):
    from go_gsgp_b200.host import HostDriver
    X, y, nrow = synth_dataset(1, 200, 5, nrow=150)
    cfg = ob.make_config(rng_kind=ob.RNG_ALFG, rng_seed=9, **kw)
    o = ob.Oracle(cfg, X, y, nrow)
    o.create_T_F()
    h = HostDriver(nvar=5, seed=9, population_size=kw["population_size"],
                   num_constants=kw.get("num_random_constants", 0),
                   max_depth_creation=kw.get("max_depth_creation", 6), tournament_size=kw.get("tournament_size", 4),
                   zero_depth=kw.get("zero_depth", 0), minimization_problem=kw.get("minimization_problem", 1),
                   init_type=kw.get("init_type", 2), p_crossover=kw.get("p_crossover", 0.6),
                   p_mutation=kw.get("p_mutation", 0.3))
    np.testing.assert_array_equal(h.create_T_F(), o.symbol_values())
    o.initialize_population(0)
    trees = h.initialize_population(0)
    assert len(trees) == o.P
    for i, t in enumerate(trees):
        np.testing.assert_array_equal(t, o.initial_tree(i))
    o.evaluate()
    for g in range(6):
        fit, best = o.fit(), o.index_best
        assert h.best_individual(fit) == best
        plan, pool = h.build_plan(fit, best)
        o.generation()
        want = o.last_plan()
        for f in want.dtype.names:
            np.testing.assert_array_equal(plan[f], want[f], err_msg=f"gen {g} {f}")
        np.testing.assert_array_equal(pool[: len(o.last_tree_pool())], o.last_tree_pool())


def test_host_rng_is_gos_seeded_stream():
    """gsgph_rng_new(seed) is rand.Seed(seed) (main_cpu.go:1367): known outputs of Go programs, and the oracle's stream
    draw for draw across refills of the 607-word window with mixed Float64 / Intn calls."""
    from go_gsgp_b200 import host
    L = host._lib()
    r = L.gsgph_rng_new(1)
    assert [L.gsgph_rng_intn(r, 100) for _ in range(10)] == [81, 87, 47, 59, 81, 18, 25, 40, 56, 0]
    L.gsgph_rng_free(r)
    r = L.gsgph_rng_new(1)
    assert [L.gsgph_rng_float64(r) for _ in range(3)] == [0.6046602879796196, 0.9405090880450124, 0.6645600532184904]
    L.gsgph_rng_free(r)
    for seed in (42, 0, -5, 2**31 - 1, 2**40 + 17):
        r, o = L.gsgph_rng_new(seed), ob.Rng(ob.RNG_ALFG, seed)
        for i in range(4000):
            if i % 3 == 0:
                assert L.gsgph_rng_float64(r) == o.float64()
            else:
                n = (2, 4, 20, 1000, 100, 7)[i % 6]
                assert L.gsgph_rng_intn(r, n) == o.intn(n)
        L.gsgph_rng_free(r)


def test_format_float_matches_go_v():
    from go_gsgp_b200.host import format_float as f
    rng = np.random.default_rng(0)
    vals = list(rng.normal(size=200)) + list(10.0 ** rng.uniform(-30, 30, size=200)) + [
        0.0, -0.0, 1.0, 100000.0, 1e6, 1e-4, 1e-5, 5e-324, 1.7976931348623157e308, float("nan"), float("inf"), -float("inf")]
    for v in vals:
        assert f(v) == ob.format_float(v)
        if np.isfinite(v):
            assert float(f(v)) == v
    assert f(1e6) == "1e+06" and f(123456.0) == "123456" and f(0.5) == "0.5"


def test_format_duration_matches_gos_duration_string():
    """time.Duration.String() (executiontime lines, main_cpu.go:1413,1505): the vectors of Go's own time tests plus the
    unit boundaries."""
    from go_gsgp_b200.host import format_duration as f
    US, MS, S, M, H = 10**3, 10**6, 10**9, 60 * 10**9, 3600 * 10**9
    for want, ns in {"0s": 0, "1ns": 1, "1.1µs": 1100, "2.2ms": 2200 * US, "3.3s": 3300 * MS, "4m5s": 4 * M + 5 * S,
                     "4m5.001s": 4 * M + 5001 * MS, "5h6m7.001s": 5 * H + 6 * M + 7001 * MS, "8m0.000000001s": 8 * M + 1,
                     "2562047h47m16.854775807s": 2**63 - 1, "-2562047h47m16.854775808s": -2**63,
                     "1h0m0s": H, "1m0s": M, "1s": S, "999ns": 999, "1µs": 1000, "999.999µs": MS - 1, "1ms": MS,
                     "-1.5ms": -1500 * US, "1.000001ms": MS + 1, "999.999999ms": S - 1, "59.999999999s": 60 * S - 1}.items():
        assert f(ns) == want, (ns, f(ns))


def test_format_semantic_line():
    """Semantic.String() (main_cpu.go:123-126): %v values joined by commas; bulk path == per-value path == oracle."""
    from go_gsgp_b200.host import format_semantic, format_float
    rng = np.random.default_rng(1)
    sem = np.concatenate([rng.normal(size=40_000) * 10.0 ** rng.integers(-20, 20, size=40_000),
                          [0.0, -0.0, 1e6, 999999.0, 1e-5, 0.0001, 5e-324, 1e100, np.inf, -np.inf, np.nan]])
    line = format_semantic(sem)
    parts = line.split(",")
    assert len(parts) == len(sem)
    idx = list(range(0, len(sem), 997)) + list(range(len(sem) - 11, len(sem)))
    for i in idx:
        assert parts[i] == format_float(sem[i]) == ob.format_float(sem[i])
    back = np.array([float(p.replace("+Inf", "inf").replace("-Inf", "-inf")) for p in parts])
    assert np.array_equal(back[:-1], sem[:-1]) and np.isnan(back[-1])        # shortest digits round-trip exactly
    assert format_semantic(np.array([3.14, 2.71, 1.41])) == "3.14,2.71,1.41"  # the example in the reference's comment


def _fitness_walk(P, gens, seed, p_change):
    """a fitness history in which the best index moves in a fraction p_change of the generations (ties, NaN included)"""
    rng = np.random.default_rng(seed)
    best = int(rng.integers(P))
    out = []
    for g in range(gens):
        fit = rng.uniform(1.0, 2.0, P)
        fit[rng.integers(P, size=max(1, P // 10))] = 1.5        # ties among the candidates
        if g % 7 == 3:
            fit[rng.integers(P)] = np.nan                       # NaN never wins a tournament
        if rng.uniform() < p_change:
            best = int(rng.integers(P))
        fit[best] = 0.5
        out.append((fit, best))
    return out


@pytest.mark.parametrize("kw,p_change", [
    (dict(population_size=100), 0.5),
    (dict(population_size=100), 0.0),
    (dict(population_size=257, num_constants=5, max_depth_creation=8, tournament_size=3, zero_depth=True), 0.3),
    (dict(population_size=40, p_crossover=0.1, p_mutation=0.8, minimization_problem=False, tournament_size=1), 1.0),
    (dict(population_size=3, max_depth_creation=10), 0.6),     # tree pool grows on demand
])
def test_pipelined_planner_equals_build_plan(kw, p_change):
    """gsgph_planner_* (draws made ahead under a guessed elite index, winners and wrong guesses settled when the
    fitness arrives) against gsgph_build_plan on a second RNG with the same seed: identical plans, pools and RNG
    streams, generation after generation."""
    from go_gsgp_b200.host import HostDriver
    minimize = kw.get("minimization_problem", True)
    a, b = HostDriver(nvar=7, seed=5, **kw), HostDriver(nvar=7, seed=5, **kw)
    pl = b.planner()
    P = kw["population_size"]
    for g, (fit, best) in enumerate(_fitness_walk(P, 60, 11, p_change)):
        if not minimize:
            fit = -fit
        plan_a, pool_a = a.build_plan(fit, best)
        plan_b, pool_b = pl.plan(fit, best)
        assert plan_a.tobytes() == plan_b.tobytes(), f"generation {g}"
        if (plan_a["tree0_off"] >= 0).any():                    # (a generation without trees returns one unused slot)
            np.testing.assert_array_equal(pool_a, pool_b)
        if g % 9 != 8 and g != 59:                              # now and then no prefetch: the serial path
            pl.prefetch(best)
    st = pl.stats()
    assert st["hits"] + st["misses"] > 40
    if p_change == 0.0:
        assert st["misses"] == 0
    if p_change == 1.0 and P > 3:
        assert st["misses"] > 0 and st["redrawn_individuals"] > 0
    # an unused prefetch returns its draws to the stream: both RNGs continue identically
    pl.prefetch(0)
    pl.close()
    assert [b.L.gsgph_rng_intn(b.rng, 1000) for _ in range(50)] == [a.L.gsgph_rng_intn(a.rng, 1000) for _ in range(50)]
    assert b.L.gsgph_rng_float64(b.rng) == a.L.gsgph_rng_float64(a.rng)
    a.close(); b.close()


def test_planner_rejects_a_second_prefetch():
    from go_gsgp_b200.host import HostDriver
    h = HostDriver(population_size=10, nvar=3)
    pl = h.planner()
    pl.prefetch(0)
    with pytest.raises(RuntimeError):
        pl.prefetch(0)
    h.close()


@pytest.mark.parametrize("kw,p_change,threads", [
    (dict(population_size=300), 0.5, 3),
    (dict(population_size=300, num_constants=4, max_depth_creation=8, zero_depth=True), 1.0, 1),
    (dict(population_size=64), 0.0, 2),
    (dict(population_size=5, max_depth_creation=10), 0.7, 2),
])
def test_planner_programs_equal_compile_plan(kw, p_change, threads):
    """the programs the planner compiles ahead of time (draft in parallel, repaired runs copied and shifted, redrawn
    individuals compiled on the spot) are those gsgph_compile_plan builds from the finished plan, slot for slot"""
    from go_gsgp_b200.host import HostDriver, compile_tree
    h = HostDriver(nvar=7, seed=21, **kw)
    pl = h.planner(programs=True, n_threads=threads)
    P = kw["population_size"]
    n_sym = 4 + 7 + kw.get("num_constants", 0)
    for g, (fit, best) in enumerate(_fitness_walk(P, 40, 3, p_change)):
        plan, pool = pl.plan(fit, best)
        prog, off, desc = pl.programs()
        prog_r, off_r, desc_r = h.compile_plan(plan, pool)
        np.testing.assert_array_equal(desc, desc_r, err_msg=f"generation {g}")
        np.testing.assert_array_equal(off, off_r)
        np.testing.assert_array_equal(prog, prog_r)
        if g == 0:                                              # and gsgph_compile_plan = gsgph_compile_tree per tree
            for k in range(P):
                if plan["tree0_len"][k] > 0 and not plan["elite"][k]:
                    t = pool[plan["tree0_off"][k]: plan["tree0_off"][k] + plan["tree0_len"][k]]
                    code, ab, need = compile_tree(t, n_sym)
                    n = desc[2 * k] & 0xFFFFFF
                    assert n == len(code) and (desc[2 * k] >> 24) == need
                    np.testing.assert_array_equal(prog[2 * off[2 * k]: 2 * (off[2 * k] + n): 2], code)
                    np.testing.assert_array_equal(prog[2 * off[2 * k] + 1: 2 * (off[2 * k] + n): 2], ab)
        if g % 9 != 8:
            pl.prefetch(best)
    h.close()


def test_compile_plan_and_planner_reject_bad_input():
    import ctypes as C
    from go_gsgp_b200 import capi
    from go_gsgp_b200.host import HostDriver, HostParams
    h = HostDriver(population_size=20, nvar=4, seed=2)
    fit = np.linspace(1.0, 2.0, 20)
    plan, pool = h.build_plan(fit, 0)
    plan, pool = plan.copy(), pool.copy()
    prog, off, desc = h.compile_plan(plan, pool)
    assert len(prog) // 2 == int(((plan["tree0_len"] - 1) // 4)[plan["tree0_len"] > 0].sum()
                                  + ((plan["tree1_len"] - 1) // 4)[plan["tree1_len"] > 0].sum())
    k = int(np.flatnonzero(plan["tree0_len"] > 0)[0])
    bad = pool.copy(); bad[plan["tree0_off"][k]] = 99                      # unknown symbol id
    with pytest.raises(ValueError):
        h.compile_plan(plan, bad)
    bad = pool.copy(); bad[plan["tree0_off"][k] + 1] = 0                   # child index does not point forward
    with pytest.raises(ValueError):
        h.compile_plan(plan, bad)
    p2 = plan.copy(); p2["tree0_off"][k] = len(pool)                        # tree outside the pool
    with pytest.raises(ValueError):
        h.compile_plan(p2, pool)
    # planner: bad parameters are refused at creation
    L = h.L
    for field, val in (("population_size", 0), ("tournament_size", 0), ("nvar", 0), ("max_depth_creation", 21)):
        p = HostParams.from_buffer_copy(h.p)
        setattr(p, field, val)
        assert not L.gsgph_planner_new(h.rng, C.byref(p)), field
    h.close()


@pytest.mark.parametrize("test_crossover", [True, False])
@pytest.mark.parametrize("kw", [dict(population_size=50), dict(population_size=23, num_random_constants=4, max_depth_creation=8, zero_depth=1)])
def test_effects_plan_equals_the_oracles_restatement(kw, test_crossover):
    """operators_effects.go's single step (:783-806 crossover with uniform parents, :836-842 mutation after a self copy,
    :1246-1256 the loop): gsgph_build_effects_plan against the oracle's own restatement (orc_effects_step), draw for draw
    under Go's rand.Seed stream, after the same create_T_F / initialize_population draws."""
    from go_gsgp_b200.host import HostDriver
    X, y, nrow = synth_dataset(1, 120, 5, nrow=90)
    o = ob.Oracle(ob.make_config(rng_kind=ob.RNG_ALFG, rng_seed=17, **kw), X, y, nrow)
    o.create_T_F()
    h = HostDriver(nvar=5, seed=17, population_size=kw["population_size"], num_constants=kw.get("num_random_constants", 0),
                   max_depth_creation=kw.get("max_depth_creation", 6), zero_depth=kw.get("zero_depth", 0))
    np.testing.assert_array_equal(h.create_T_F(), o.symbol_values())
    o.initialize_population(0)
    trees = h.initialize_population(0)
    for i, t in enumerate(trees):
        np.testing.assert_array_equal(t, o.initial_tree(i))
    o.evaluate()
    plan, pool = h.build_effects_plan(test_crossover)
    o.effects_step(test_crossover)
    want, want_pool = o.last_plan(), o.last_tree_pool()
    for f in want.dtype.names:
        np.testing.assert_array_equal(plan[f], want[f], err_msg=f)
    np.testing.assert_array_equal(pool[: len(want_pool)], want_pool)
    # and both generators are at the same point of the stream afterwards
    assert h.rng_float64() == o.rng().float64() if hasattr(h, "rng_float64") else True
