"""Builds an (oracle, engine) pair on the same inputs for the GPU parity tests."""
import numpy as np

from oracle import binding as ob
from go_gsgp_b200 import Engine, capi


def make_pair(X, y, nrow, *, population_size, rng_kind=ob.RNG_ALFG, seed=1, error_measure=ob.RMSE,
              max_depth_creation=6, tournament_size=4, p_crossover=0.6, p_mutation=0.3,
              num_random_constants=0, minimization_problem=1, zero_depth=0, init_type=2,
              seeds=None, flags=0, workspace_bytes=0, sym_override=None, n_workers=1, use_linear_scaling=0):
    cfg = ob.make_config(population_size=population_size, rng_kind=rng_kind, rng_seed=seed,
                         error_measure=error_measure, max_depth_creation=max_depth_creation,
                         tournament_size=tournament_size, p_crossover=p_crossover, p_mutation=p_mutation,
                         num_random_constants=num_random_constants, minimization_problem=minimization_problem,
                         zero_depth=zero_depth, init_type=init_type, n_workers=n_workers,
                         use_linear_scaling=use_linear_scaling)
    o = ob.Oracle(cfg, X, y, nrow)
    o.create_T_F()
    N, V = X.shape
    e = Engine(population_size=population_size, nvar=V, nrow=nrow, nrow_test=N - nrow,
               max_depth_creation=max_depth_creation, tournament_size=tournament_size,
               zero_depth=bool(zero_depth), minimization_problem=bool(minimization_problem),
               error_measure=error_measure, num_constants=num_random_constants,
               p_crossover=p_crossover, p_mutation=p_mutation, rng_seed=seed,
               rng_mode=capi.RNG_DEVICE_PHILOX if rng_kind == ob.RNG_PHILOX else capi.RNG_HOST_REPLAY,
               flags=flags, workspace_bytes=workspace_bytes, use_linear_scaling=bool(use_linear_scaling))
    e.set_dataset(X, y)
    sv = o.symbol_values() if sym_override is None else np.asarray(sym_override, np.float64)
    e.set_constants(sv)
    seeds = seeds or []
    for i, s in enumerate(seeds):
        o.seed_semantic(i, s)
        e.seed_semantic(i, s)
    o.initialize_population(len(seeds))
    trees = [o.initial_tree(i) for i in range(len(seeds), population_size)]
    if trees:
        e.eval_initial(len(seeds), trees)
    o.evaluate()
    fit, fit_test, best = e.fitness_all()
    return o, e, (fit, fit_test, best)
