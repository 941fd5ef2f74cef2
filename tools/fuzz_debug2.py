import sys
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import binding as ob
from tests.pair import make_pair
from tests.util import synth_dataset, rel_err

def truth_ls(y, s, nrow, measure):
    yl, sl = y.astype(np.longdouble), s.astype(np.longdouble)
    t, o = yl[:nrow], sl[:nrow]
    ob_, tb = o.mean(), t.mean()
    den = ((o - ob_) ** 2).sum(); num = ((t - tb) * (o - ob_)).sum()
    b = num / den if den != 0 else 0; a = tb - b * ob_
    r = t - (a + b * o)
    d = {0: (r ** 2).mean(), 1: np.sqrt((r ** 2).mean()), 2: np.abs(r).mean(), 3: (np.abs(r) / t).mean()}[measure]
    return float(d), float(a), float(b), float(den)

for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + case)
    v = int(rng.integers(5, 40)); nrow = int(rng.integers(1, 3000)); nrow_test = int(rng.integers(0, 900)); P = int(rng.integers(2, 90))
    X, y, _ = synth_dataset(7, nrow + nrow_test, v, nrow=nrow)
    if case % 3 == 0: y = y * 50.0 - 7.0
    kw = dict(population_size=P, seed=int(rng.integers(1, 1 << 30)), error_measure=int(rng.choice([ob.MSE, ob.RMSE, ob.MAE, ob.MRE])),
              max_depth_creation=int(rng.integers(1, 10)), tournament_size=int(rng.integers(1, 9)),
              num_random_constants=int(rng.choice([0, 0, 3, 17])), zero_depth=int(rng.integers(0, 2)),
              minimization_problem=int(rng.integers(0, 2)), init_type=int(rng.integers(0, 3)),
              use_linear_scaling=int(rng.integers(0, 2)), n_workers=int(rng.choice([1, 1, 4])), p_crossover=float(rng.choice([0.0, 0.3, 0.6, 1.0])))
    kw["p_mutation"] = float(min(1.0 - kw["p_crossover"], rng.choice([0.0, 0.3, 0.4, 1.0])))
    if not kw['use_linear_scaling'] or kw['n_workers'] != 1 or nrow < 10: continue
    o, e, (fit, fit_test, best) = make_pair(X, y, nrow, **kw)
    for g in range(3):
        o.generation()
        fit, fit_test, best = e.generation(o.last_plan(), o.last_tree_pool())
        err = rel_err(fit, o.fit())
        if err.max() > 1e-9:
            k = int(np.argmax(err)); s = o.semantic(k)
            tr = truth_ls(y, s, nrow, kw['error_measure'])
            f, a, b = o.fitness_train_ls(s[:nrow])
            sg = e.get_semantic(k)
            print(case, 'gen', g, 'ind', k, 'measure', kw['error_measure'], 'gpu', repr(fit[k]), 'oracle', repr(o.fit()[k]), 'truth', tr[0], '| oracle a,b', a, b, 'truth a,b,den', tr[1:], '| sem mean', s[:nrow].mean(), 'std', s[:nrow].std(), 'max|s|', np.abs(s).max(), 'sem relerr', rel_err(sg, s).max(), 'plan', o.last_plan()[k])
            break
    e.close(); o.close()
