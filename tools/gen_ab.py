#!/usr/bin/env python
"""A/B timing of the fused generation kernel's variants (GSGP_GEN_WARPS / GSGP_GEN_WAVES) in one process.

    python tools/gen_ab.py [--workload cfg2] [--steps 30] [--variants 16,20,24] [--pop N]

For every variant: a fresh device-mode engine on the same synthetic workload, W warm-up generations, K generations
timed with the library's own per-launch CUDA events (gsgp_profile_read: the generation kernel's launches only) and
with one event pair around the whole run.  All variants run the same arithmetic in the same order, so their final
fitness vectors must be bit-identical: the script checks that against the first variant.
Prints one JSON line per variant.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--variants", default="20,24,28")
    ap.add_argument("--pop", type=int, default=0)
    ap.add_argument("--cases", type=int, default=0)
    args = ap.parse_args()

    import torch
    import bench
    from go_gsgp_b200 import Engine, capi
    from go_gsgp_b200.host import HostDriver

    w = dict(bench.WORKLOADS[args.workload])
    if args.pop:
        w["pop"] = args.pop
    if args.cases:
        w["cases_per_gpu"] = args.cases
    P, V, n = w["pop"], w["nvar"], w["cases_per_gpu"]
    nrow = int(0.8 * n)
    X, y = bench.synth(w["cfg_id"], n, V)
    host = HostDriver(population_size=P, nvar=V, max_depth_creation=w["depth"], p_crossover=w["p_xo"],
                      p_mutation=w["p_mut"], seed=1)
    sym = host.create_T_F()
    init_trees = host.initialize_population(0)
    ref_fit = None
    for var in args.variants.split(","):
        # a variant is warps[:waves], e.g. 28, 24, 28:12
        parts = var.split(":")
        os.environ["GSGP_GEN_WARPS"] = parts[0]
        if len(parts) > 1 and parts[1]:
            os.environ["GSGP_GEN_WAVES"] = parts[1]
        else:
            os.environ.pop("GSGP_GEN_WAVES", None)
        os.environ["GSGP_PIPELINE"] = "fused"
        e = Engine(rng_mode=capi.RNG_DEVICE_PHILOX, population_size=P, nvar=V, nrow=nrow, nrow_test=n - nrow,
                   max_depth_creation=w["depth"], error_measure=capi.RMSE, p_crossover=w["p_xo"], p_mutation=w["p_mut"],
                   rng_seed=1)
        e.set_dataset(X, y)
        e.set_constants(sym)
        e.eval_initial(0, init_trees)
        e.fitness_all()
        stream = torch.cuda.ExternalStream(e.stream)
        e.run_generations(args.warmup)
        e.sync()
        e.profile_enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        e.run_generations(args.steps)
        ev1.record(stream)
        e.sync()
        ms = ev0.elapsed_time(ev1)
        n_gs, ms_gs, _ = e.profile_read(capi.K_GSOP)
        e.profile_enable(False)
        fit, fit_test, best = e.get_fitness()
        same = None
        if ref_fit is None:
            ref_fit = (fit.copy(), fit_test.copy(), best)
        else:
            same = bool(np.array_equal(fit, ref_fit[0], equal_nan=True) and np.array_equal(fit_test, ref_fit[1], equal_nan=True)
                        and best == ref_fit[2])
        e.close()
        print(json.dumps({"variant": var, "workload": args.workload, "pop": P, "cases": n, "ms_per_generation": ms / args.steps,
                          "gen_kernel_ms_per_launch": ms_gs / max(1, n_gs), "launches": n_gs,
                          "updates_per_s": P * n * args.steps / (ms * 1e-3), "bit_identical_to_first": same,
                          "best_fit": float(fit[best])}), flush=True)


if __name__ == "__main__":
    main()
