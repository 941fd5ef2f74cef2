import sys, os
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import binding as ob
from tests.pair import make_pair
from tests.util import synth_dataset, rel_err
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
  if not kw['use_linear_scaling']: continue
  o, e, (fit, fit_test, best) = make_pair(X, y, nrow, **kw)
  err = rel_err(fit, o.fit())
  if err.max() <= 1e-9: e.close(); o.close(); continue
  print(case, dict(v=v, nrow=nrow, nrow_test=nrow_test), kw)
  bad = np.argsort(-err)[:4]
  for k in bad:
      s = o.semantic(k)[:nrow]
      f, a, b = o.fitness_train_ls(s) if kw["use_linear_scaling"] else (o.fit()[k], 0, 0)
      print(f"ind {k}: gpu fit {fit[k]!r} oracle {o.fit()[k]!r} err {err[k]:.3e} | sem min {s.min()!r} max {s.max()!r} std {s.std():.3e} | oracle a {a!r} b {b!r} | tree {list(o.initial_tree(k))[:9]}")
  e.close(); o.close()
