#!/usr/bin/env python
"""Generates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors for this path and cannot be run here (no Go toolchain), so
these fixtures pin the ORACLE (and through it the CUDA path) against drift; they do not pin parity with go-gsgp.
Contents: BASELINE config 1 shape (5 vars x 1000/250 cases, pop 100), Philox streams (seed 1), 6 generations:
best-of-generation history, the full fitness vectors after the last generation, the first generation's decisions, and
the hand-derived known answers of SURVEY Appendix B evaluated by the oracle.  Floats are stored as hex (exact).

cfg1_go_seed1.json: the same shape under Go's own math/rand stream for `rand.Seed(1)` (`-seed 1`; the flag's default is the clock,
main_cpu.go:172,1367; the stream is pinned to real Go outputs, tools/gen_go_rng_cooked.py) with sequential sums
(`-workers 1`): initial trees, first-generation plan and tree pool, best-of-generation history.  The product's host half
(gsgph_rng / gsgph_initialize_population / gsgph_build_plan) is tested against it without the oracle in the loop."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob                      # noqa: E402
from tests.util import synth_dataset, toy_fixture     # noqa: E402


def hexes(a):
    return [float(v).hex() for v in np.asarray(a, np.float64).ravel()]


def main():
    out = {}
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    for name, extra in (("cfg1_rmse", dict(error_measure=ob.RMSE)),
                        ("cfg1_mae_linsc", dict(error_measure=ob.MAE, use_linear_scaling=1, num_random_constants=3))):
        cfg = ob.make_config(population_size=100, rng_kind=ob.RNG_PHILOX, rng_seed=1, **extra)
        o = ob.Oracle(cfg, X, y, nrow)
        o.create_T_F(); o.initialize_population(0); o.evaluate()
        hist = [(o.index_best, o.fit()[o.index_best], o.fit_test()[o.index_best])]
        plan1 = None
        for g in range(6):
            o.generation()
            if g == 0:
                p = o.last_plan()
                plan1 = {k: p[k].tolist() for k in ("op", "elite", "p1", "p2", "tree0_len", "tree1_len")}
                plan1["ms"] = hexes(p["ms"])
            hist.append((o.index_best, o.fit()[o.index_best], o.fit_test()[o.index_best]))
        out[name] = {"config": {k: (v if not isinstance(v, float) else v) for k, v in extra.items()},
                     "best_index": [h[0] for h in hist], "best_fit": hexes([h[1] for h in hist]),
                     "best_fit_test": hexes([h[2] for h in hist]), "fit": hexes(o.fit()), "fit_test": hexes(o.fit_test()),
                     "semantic_of_best_head": hexes(o.semantic(o.index_best)[:16]), "plan_generation_1": plan1}
        o.close()
    # Appendix B known answers on the 4-row fixture of main_test.go:107-112
    Xt, yt, nt = toy_fixture()
    o = ob.Oracle(ob.make_config(population_size=2), Xt, yt, nt)
    o.create_T_F()
    out["toy"] = {"tree_plus_x0_times_x1_x0": hexes(o.semantic_evaluate_array(np.array([0, 3, 4, 4, 2, 7, 8, 5, 4], np.int32))),
                  "tree_x0_div_x1_minus_x1": hexes(o.semantic_evaluate_array(np.array([3, 3, 4, 4, 1, 7, 8, 5, 5], np.int32)))}
    o.close()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg1_philox_seed1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", f.name)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.int32).tobytes()).hexdigest()


def main_go_stream():
    X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
    cfg = ob.make_config(population_size=100, rng_kind=ob.RNG_ALFG, rng_seed=1, error_measure=ob.RMSE, n_workers=1)
    o = ob.Oracle(cfg, X, y, nrow)
    o.create_T_F(); o.initialize_population(0)
    trees = [o.initial_tree(i) for i in range(100)]
    o.evaluate()
    out = {"config": {"error_measure": ob.RMSE}, "initial_tree_len": [len(t) for t in trees],
           "initial_trees_sha256": sha(np.concatenate(trees)), "initial_tree_0": trees[0].tolist(),
           "initial_fit": hexes(o.fit()), "initial_fit_test": hexes(o.fit_test())}
    hist = [(o.index_best, o.fit()[o.index_best], o.fit_test()[o.index_best])]
    for g in range(6):
        o.generation()
        if g == 0:
            p = o.last_plan()
            out["plan_generation_1"] = {k: p[k].tolist() for k in ("op", "elite", "p1", "p2", "tree0_off", "tree0_len", "tree1_off", "tree1_len")}
            out["plan_generation_1"]["ms"] = hexes(p["ms"])
            out["plan_generation_1"]["tree_pool_sha256"] = sha(o.last_tree_pool())
            out["plan_generation_1"]["tree_pool_len"] = int(len(o.last_tree_pool()))
        hist.append((o.index_best, o.fit()[o.index_best], o.fit_test()[o.index_best]))
    out.update({"best_index": [h[0] for h in hist], "best_fit": hexes([h[1] for h in hist]),
                "best_fit_test": hexes([h[2] for h in hist]), "fit": hexes(o.fit()), "fit_test": hexes(o.fit_test()),
                "semantic_of_best_head": hexes(o.semantic(o.index_best)[:16])})
    o.close()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg1_go_seed1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", f.name)


if __name__ == "__main__":
    main()
    main_go_stream()
