"""Committed golden fixtures (tests/golden/, generated from the oracle by make_golden.py).  They pin the oracle -- and,
on a GPU, the CUDA path -- against drift; the reference itself ships no vectors for this path (parity unpinned)."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from tests.util import synth_dataset, toy_fixture, rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfg1_philox_seed1.json")))


GO = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfg1_go_seed1.json")))


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, np.int32).tobytes()).hexdigest()


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


@pytest.fixture(scope="module", autouse=True)
def _built():
    ob.build()


@pytest.mark.parametrize("name", ["cfg1_rmse", "cfg1_mae_linsc"])
def test_oracle_reproduces_golden_run(name):
    g = GOLD[name]
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    o = ob.Oracle(ob.make_config(population_size=100, rng_kind=ob.RNG_PHILOX, rng_seed=1, **g["config"]), X, y, nrow)
    o.create_T_F(); o.initialize_population(0); o.evaluate()
    best, bf, bt = [o.index_best], [o.fit()[o.index_best]], [o.fit_test()[o.index_best]]
    for gen in range(6):
        o.generation()
        if gen == 0:
            p = o.last_plan()
            for k in ("op", "elite", "p1", "p2", "tree0_len", "tree1_len"):
                assert p[k].tolist() == g["plan_generation_1"][k], k
            assert np.array_equal(p["ms"], unhex(g["plan_generation_1"]["ms"]))
        best.append(o.index_best); bf.append(o.fit()[o.index_best]); bt.append(o.fit_test()[o.index_best])
    assert best == g["best_index"]
    assert np.all(rel_err(bf, unhex(g["best_fit"])) <= 1e-12) and np.all(rel_err(bt, unhex(g["best_fit_test"])) <= 1e-12)
    assert np.all(rel_err(o.fit(), unhex(g["fit"])) <= 1e-12) and np.all(rel_err(o.fit_test(), unhex(g["fit_test"])) <= 1e-12)
    assert np.all(rel_err(o.semantic(o.index_best)[:16], unhex(g["semantic_of_best_head"])) <= 1e-12)
    o.close()


def test_oracle_reproduces_toy_known_answers():
    X, y, nrow = toy_fixture()
    o = ob.Oracle(ob.make_config(population_size=2), X, y, nrow)
    o.create_T_F()
    a = o.semantic_evaluate_array(np.array([0, 3, 4, 4, 2, 7, 8, 5, 4], np.int32))
    assert np.array_equal(a, unhex(GOLD["toy"]["tree_plus_x0_times_x1_x0"]))
    assert np.array_equal(a, X[:, 0] + X[:, 1] * X[:, 0])                   # SURVEY Appendix B, by hand
    b = o.semantic_evaluate_array(np.array([3, 3, 4, 4, 1, 7, 8, 5, 5], np.int32))
    assert np.array_equal(b, unhex(GOLD["toy"]["tree_x0_div_x1_minus_x1"])) and np.all(b == 1.0)   # protected division
    o.close()


def test_oracle_reproduces_go_stream_golden():
    """BASELINE config 1 under Go's own rand.Seed(1) stream (main_cpu.go:1367), sequential sums (-workers 1)."""
    g = GO
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    o = ob.Oracle(ob.make_config(population_size=100, rng_kind=ob.RNG_ALFG, rng_seed=1, n_workers=1, **g["config"]), X, y, nrow)
    o.create_T_F(); o.initialize_population(0)
    trees = [o.initial_tree(i) for i in range(100)]
    assert [len(t) for t in trees] == g["initial_tree_len"] and sha(np.concatenate(trees)) == g["initial_trees_sha256"]
    o.evaluate()
    assert np.all(rel_err(o.fit(), unhex(g["initial_fit"])) <= 1e-12)
    best, bf = [o.index_best], [o.fit()[o.index_best]]
    for gen in range(6):
        o.generation()
        if gen == 0:
            assert sha(o.last_tree_pool()) == g["plan_generation_1"]["tree_pool_sha256"]
        best.append(o.index_best); bf.append(o.fit()[o.index_best])
    assert best == g["best_index"] and np.all(rel_err(bf, unhex(g["best_fit"])) <= 1e-12)
    assert np.all(rel_err(o.fit(), unhex(g["fit"])) <= 1e-12) and np.all(rel_err(o.fit_test(), unhex(g["fit_test"])) <= 1e-12)
    o.close()


def _host_driver_seed1():
    from go_gsgp_b200.host import HostDriver
    return HostDriver(population_size=100, nvar=5, seed=1)       # the reference's default flags (main_cpu.go:157-169) with -seed 1


def test_host_half_reproduces_go_stream_golden():
    """The product's host half alone (no oracle in the loop): gsgph_rng_new(1) + gsgph_initialize_population +
    gsgph_build_plan give the fixture's initial trees and -- fed the fixture's fitness -- its first-generation plan and
    tree pool, i.e. the draws main_cpu.go makes from rand.Seed(1)."""
    g, pg = GO, GO["plan_generation_1"]
    h = _host_driver_seed1()
    h.create_T_F()
    trees = h.initialize_population(0)
    assert [len(t) for t in trees] == g["initial_tree_len"]
    assert trees[0].tolist() == g["initial_tree_0"] and sha(np.concatenate(trees)) == g["initial_trees_sha256"]
    plan, pool = h.build_plan(unhex(g["initial_fit"]), g["best_index"][0])
    for k in ("op", "elite", "p1", "p2", "tree0_off", "tree0_len", "tree1_off", "tree1_len"):
        assert plan[k].tolist() == pg[k], k
    assert np.array_equal(plan["ms"], unhex(pg["ms"]))
    assert len(pool) == pg["tree_pool_len"] and sha(pool) == pg["tree_pool_sha256"]
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipelined", [False, True])
def test_cuda_path_reproduces_go_stream_golden(pipelined):
    """Host half + CUDA path, no oracle in the loop: six generations of config 1 from rand.Seed(1) land on the fixture's
    best-of-generation history and final fitness (1e-9: the fixture's sums are sequential, the device's are trees)."""
    from go_gsgp_b200 import Engine, capi
    g = GO
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    h = _host_driver_seed1()
    sym, trees = h.create_T_F(), h.initialize_population(0)
    e = Engine(population_size=100, nvar=5, nrow=nrow, nrow_test=250, error_measure=g["config"]["error_measure"])
    e.set_dataset(X, y); e.set_constants(sym); e.eval_initial(0, trees)
    fit, fit_test, best = e.fitness_all()
    assert np.all(rel_err(fit, unhex(g["initial_fit"])) <= 1e-9) and np.all(rel_err(fit_test, unhex(g["initial_fit_test"])) <= 1e-9)
    pl = h.planner() if pipelined else None
    bi, bf, bt = [best], [fit[best]], [fit_test[best]]
    for gen in range(6):
        if pl is None:
            plan, pool = h.build_plan(fit, best)
        else:
            plan, pool = pl.plan(fit, best)
            pl.prefetch(best)
        if gen == 0:
            assert plan["p1"].tolist() == g["plan_generation_1"]["p1"] and sha(pool) == g["plan_generation_1"]["tree_pool_sha256"]
        fit, fit_test, best = e.generation(plan, pool)
        bi.append(best); bf.append(fit[best]); bt.append(fit_test[best])
    assert bi == g["best_index"]
    assert np.all(rel_err(bf, unhex(g["best_fit"])) <= 1e-9) and np.all(rel_err(bt, unhex(g["best_fit_test"])) <= 1e-9)
    assert np.all(rel_err(fit, unhex(g["fit"])) <= 1e-9) and np.all(rel_err(fit_test, unhex(g["fit_test"])) <= 1e-9)
    assert np.all(rel_err(e.get_semantic(best)[:16], unhex(g["semantic_of_best_head"])) <= 1e-9)
    e.close(); h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1_rmse", "cfg1_mae_linsc"])
def test_cuda_path_reproduces_golden_run(name):
    from go_gsgp_b200 import Engine, capi
    from go_gsgp_b200.host import HostDriver
    g = GOLD[name]
    cfgk = g["config"]
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    # the initial population comes from the oracle's host half (trees are RNG-order data, not device work)
    o = ob.Oracle(ob.make_config(population_size=100, rng_kind=ob.RNG_PHILOX, rng_seed=1, **cfgk), X, y, nrow)
    o.create_T_F(); o.initialize_population(0)
    e = Engine(population_size=100, nvar=5, nrow=nrow, nrow_test=250, error_measure=cfgk["error_measure"],
               num_constants=cfgk.get("num_random_constants", 0), rng_seed=1, rng_mode=capi.RNG_DEVICE_PHILOX,
               use_linear_scaling=bool(cfgk.get("use_linear_scaling", 0)))
    e.set_dataset(X, y)
    e.set_constants(o.symbol_values())
    e.eval_initial(0, [o.initial_tree(i) for i in range(100)])
    e.fitness_all()
    e.run_generations(1)
    p = e.get_plan()
    live = np.array(g["plan_generation_1"]["elite"]) == 0
    for k in ("op", "elite"):
        assert p[k].tolist() == g["plan_generation_1"][k], k
    for k in ("p1", "p2", "tree0_len", "tree1_len"):
        assert np.array_equal(p[k][live], np.array(g["plan_generation_1"][k])[live]), k
    assert np.array_equal(p["ms"][live], unhex(g["plan_generation_1"]["ms"])[live])
    e.run_generations(5)
    bf, bt, bi = e.get_history(0, 7)
    assert bi.tolist() == g["best_index"]
    assert np.all(rel_err(bf, unhex(g["best_fit"])) <= 1e-9) and np.all(rel_err(bt, unhex(g["best_fit_test"])) <= 1e-9)
    fit, fit_test, best = e.get_fitness()
    assert np.all(rel_err(fit, unhex(g["fit"])) <= 1e-9) and np.all(rel_err(fit_test, unhex(g["fit_test"])) <= 1e-9)
    assert np.all(rel_err(e.get_semantic(best)[:16], unhex(g["semantic_of_best_head"])) <= 1e-9)
    e.close(); o.close()
