"""Committed golden fixtures (tests/golden/, generated from the oracle by make_golden.py).  They pin the oracle -- and,
on a GPU, the CUDA path -- against drift; the reference itself ships no vectors for this path (parity unpinned)."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from tests.util import synth_dataset, toy_fixture, rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfg1_philox_seed1.json")))


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
