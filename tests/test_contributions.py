"""Contributions (main_cpu.go:128-149,857,863-873,885,909,1196,1390-1397): gsgph_contrib_* against the reference's rule
restated with Python's arbitrary-precision integers."""
import time

import numpy as np
import pytest

from go_gsgp_b200 import capi


def _model_new(P, n_seeds):                                    # init_tables + :1390-1397
    return [[1 if m == (i + 1 if i < n_seeds else 0) else 0 for m in range(n_seeds + 1)] for i in range(P)]


def _model_apply(cur, plan):
    new = []
    for k, e in enumerate(plan):
        if e["elite"]:                                         # the elite keeps its own vector (:847,857,909)
            new.append(list(cur[k]))
        elif e["op"] == capi.OP_XO:                            # merge_contribs :885
            new.append([a + b for a, b in zip(cur[e["p1"]], cur[e["p2"]])])
        else:                                                  # copy_contrib :857
            new.append(list(cur[e["p1"]]))
    return new


def _random_plan(rng, P, best, p_xo=0.6, p_mut=0.3):
    plan = np.zeros(P, capi.PLAN_DTYPE)
    u = rng.random(P)
    plan["op"] = np.where(u < p_xo, capi.OP_XO, np.where(u < p_xo + p_mut, capi.OP_MUT, capi.OP_REPR))
    plan["p1"], plan["p2"] = rng.integers(0, P, P), rng.integers(0, P, P)
    plan["p2"][plan["op"] != capi.OP_XO] = -1
    plan["elite"][best] = 1
    plan["p1"][best], plan["p2"][best] = (best, -1) if rng.random() < 0.5 else (-7, 10**6)   # device-mode plans leave them undefined
    return plan


@pytest.mark.parametrize("P,n_seeds,gens,threads", [(30, 4, 330, 1), (64, 0, 200, 3), (500, 40, 70, 0), (1, 1, 5, 1), (7, 7, 140, 2)])
def test_contributions_equal_big_integer_model(P, n_seeds, gens, threads):
    from go_gsgp_b200.host import Contributions
    rng = np.random.default_rng(P + gens)
    c, model = Contributions(P, n_seeds), _model_new(P, n_seeds)
    assert c.num_models == n_seeds + 1 and c.limbs == 1
    fmt = lambda row: ",".join(str(v) for v in row)            # Contribution.String() :141-146
    for k in range(P):
        assert c.format(k) == fmt(model[k])
    for g in range(gens):
        plan = _random_plan(rng, P, int(rng.integers(P)), p_xo=0.9 if g % 2 else 0.6)
        c.apply(plan, threads)
        model = _model_apply(model, plan)
        for k in ([0, P - 1, int(rng.integers(P))] if g % 10 else range(P)):
            assert c.format(k) == fmt(model[k]), (g, k)
    top = max(max(r) for r in model)
    assert c.limbs >= (top.bit_length() + 63) // 64            # the limb arrays widened as the values grew
    if P > 1 and gens >= 140:
        assert top.bit_length() > 64
    c.close()


def test_contributions_reject_bad_plans_and_sizes():
    from go_gsgp_b200.host import Contributions
    with pytest.raises(ValueError):
        Contributions(0, 0)
    with pytest.raises(ValueError):
        Contributions(3, 4)
    c = Contributions(5, 2)
    plan = _random_plan(np.random.default_rng(0), 5, 0)
    plan["op"][1], plan["p1"][1], plan["p2"][1] = capi.OP_XO, 2, 5
    with pytest.raises(ValueError):
        c.apply(plan)
    plan["p2"][1], plan["p1"][3] = 4, -1
    with pytest.raises(ValueError):
        c.apply(plan)
    assert c.format(0) == "0,1,0" and c.format(4) == "1,0,0"   # untouched by the rejected plans
    c.close()


def test_contributions_config3_shape_is_a_streaming_pass():
    """BASELINE config 3 (pop 2,000, 500 seeded models): a generation of bookkeeping stays far below the 12.6 ms the
    device needs for that generation, also once the values need several limbs."""
    from go_gsgp_b200.host import Contributions
    P, S = 2000, 500
    rng = np.random.default_rng(3)
    c = Contributions(P, S)
    times = []
    for g in range(150):
        plan = _random_plan(rng, P, 0, p_xo=1.0)
        t0 = time.perf_counter(); c.apply(plan); times.append(time.perf_counter() - t0)
    assert c.limbs >= 2
    assert np.median(times[-20:]) < 0.05, np.median(times[-20:])
    total = sum(int(v) for v in c.format(17).split(","))
    assert 2 ** 64 < total <= 2 ** 150                                  # a crossover adds the parents' totals
    c.close()
