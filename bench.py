#!/usr/bin/env python
"""bench.py -- individual-case semantic updates/s of the go-gsgp generation loop on B200.

A "step" is one generation: one pass of the hot path (random-tree evaluation -> GS crossover /
mutation / reproduction on the train+test semantics -> per-individual fitness -> selection) over
the whole population.  One update = one fp64 element child[individual][case] produced.

  value   device-RNG mode, everything resident in HBM (selection and random trees drawn on the
          device): K generations timed with CUDA events on the library's stream, max over ranks.
  e2e     host-replay mode through the reference-facing C-ABI call gsgp_generation() with HOST
          buffers: every step the host builds the plan (tournaments on the previous generation's
          fitness, random trees), the call uploads plan + programs, runs, and reads the fitness back.
  roofline  the GS-operator+fitness kernel: algorithmic bytes (32 B per crossover/mutation update,
          16 B per copied update) / CUDA-event time of its launches, against MEASURED_PEAKS.json.
  cpu_baseline  the C oracle (a restatement of main_cpu.go, NOT go-gsgp itself) on the host cores.

N>1: fitness cases are sharded; each rank holds `cases-per-gpu` cases of every individual (weak
scaling), the only exchange is the 2*pop partial-sum all-gather (NCCL) per generation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 20 vars x 100k cases, pop 1000, crossover+mutation, 1 B200
    "cfg2": dict(cfg_id=2, nvar=20, cases_per_gpu=100_000, pop=1000, depth=6, p_xo=0.6, p_mut=0.3,
                 desc="synthetic 20 vars x 100k cases (80k train/20k test) per GPU, pop 1000, depth 6, p_xo 0.6 p_mut 0.3"),
    # BASELINE.json configs[0]: the CPU-runnable case (latency-bound on a GPU)
    "cfg1": dict(cfg_id=1, nvar=5, cases_per_gpu=1250, pop=100, depth=6, p_xo=0.6, p_mut=0.3,
                 desc="synthetic 5 vars x 1000/250 cases, pop 100"),
    # BASELINE.json configs[2]: semantic feeding: 500 precomputed semantics (seeded individuals) x 1M cases, pop 2000.
    # The seeds are generated in memory (y + N(0, 0.05 + 0.001 k)) instead of 500 text files.
    "cfg3": dict(cfg_id=3, nvar=20, cases_per_gpu=1_000_000, pop=2000, depth=6, p_xo=0.6, p_mut=0.3, n_seeds=500,
                 desc="semantic feeding: 500 seeded semantics x 1M cases (20 vars), pop 2000, depth 6"),
    # BASELINE.json configs[3]: deep trees, tree-eval bound (pop reduced to fit the default run time)
    "cfg4": dict(cfg_id=4, nvar=20, cases_per_gpu=1_000_000, pop=5000, depth=8, p_xo=0.0, p_mut=1.0,
                 desc="synthetic 20 vars x 1M cases per GPU, pop 5000, depth 8, p_mut 1.0"),
    # BASELINE.json configs[4] per-GPU share at 8 GPUs with a population that fits one GPU
    "cfg5": dict(cfg_id=5, nvar=20, cases_per_gpu=1_250_000, pop=1000, depth=6, p_xo=0.6, p_mut=0.3,
                 desc="synthetic 20 vars x 1.25M cases per GPU (10M at 8 GPUs), pop 1000"),
    # BASELINE.json configs[4] at its full population: 10M cases x pop 10,000 at 8 GPUs = 100 GB of semantics per GPU;
    # two buffers do not fit 180 GB, so the library switches to the single-buffer store (case-block scratch)
    "cfg5full": dict(cfg_id=5, nvar=20, cases_per_gpu=1_250_000, pop=10_000, depth=6, p_xo=0.6, p_mut=0.3,
                     desc="synthetic 20 vars x 1.25M cases per GPU (10M at 8 GPUs), pop 10,000, single-buffer store"),
    # BASELINE.json configs[4], strong scaling (north_star: ">= 6x at 8 GPUs on the 10M-case config"): 10M cases in TOTAL at
    # every GPU count, with a population that fits one GPU (pop 1000 x 10M cases = 80 GB of semantics: the single-buffer
    # store at N = 1, two buffers from N = 2 on).  --scaling strong is implied.
    "cfg5s": dict(cfg_id=5, nvar=20, cases_per_gpu=10_000_000, pop=1000, depth=6, p_xo=0.6, p_mut=0.3, strong=True,
                  desc="synthetic 20 vars x 10M cases in total (strong scaling: 10M / N per GPU), pop 1000"),
}


def synth(cfg_id, n, v):
    rng = np.random.Generator(np.random.PCG64(20261019 + cfg_id))
    X = rng.uniform(-1.0, 1.0, size=(n, v))
    y = X[:, 0] * X[:, 1] + 0.5 * X[:, 2] - X[:, 3] + 0.25 * X[:, 4] ** 2 + rng.normal(0.0, 0.01, size=n)
    return X, y


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ref_arm_pop(w, n_cases):
    """population subsample of the reference (CPU) arm: ~2e7 updates per generation, at the full case count"""
    return max(8, min(w["pop"], int(2e7 // n_cases)))


def config_dict(workload, w, world, scaling):
    """`config` of the JSON line: the same dict in both arms (the driver compares them)"""
    n_global = w["cases_per_gpu"] if scaling == "strong" else w["cases_per_gpu"] * world
    n_local = n_global // world if scaling == "strong" else w["cases_per_gpu"]
    return {"workload": f"{workload}: {w['desc']}", "global_cases": n_global, "pop": w["pop"], "scaling": scaling,
            "l2": "inputs larger than L2 (semantic store %.2f GB per buffer per GPU)" % (w["pop"] * n_local * 8 / 1e9),
            "reference_arm": "--impl reference runs the CPU restatement of main_cpu.go at %d cases with a population subsample of "
                             "%d and reports updates/s (its cost is linear in the population)" % (n_local, ref_arm_pop(w, n_local))}


# ------------------------------------------------------------------------------------------------
def oracle_arm(w, n_cases, pop, gens, warm, workers):
    """Times the C restatement of main_cpu.go (the oracle) on the host cores: `gens` generations
    at the full case count with a population of `pop`.  Returns updates/s and seconds."""
    from oracle import binding as ob
    X, y = synth(w["cfg_id"], n_cases, w["nvar"])
    nrow = int(0.8 * n_cases)
    cfg = ob.make_config(population_size=pop, rng_kind=ob.RNG_ALFG, rng_seed=1, error_measure=ob.RMSE,
                         max_depth_creation=w["depth"], p_crossover=w["p_xo"], p_mutation=w["p_mut"],
                         n_workers=workers)
    o = ob.Oracle(cfg, X, y, nrow)
    o.create_T_F(); o.initialize_population(0); o.evaluate()
    for _ in range(warm):
        o.generation()
    t0 = time.perf_counter()
    for _ in range(gens):
        o.generation()
    dt = time.perf_counter() - t0
    o.close()
    return pop * n_cases * gens / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  go-gsgp cannot be built here
    (no Go toolchain), so this is the oracle port of main_cpu.go with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    workers = os.cpu_count() or 1
    world = max(1, args.gpus)
    scaling = args.scaling or ("strong" if w.get("strong") else "weak")
    n_cases = w["cases_per_gpu"] // world if scaling == "strong" else w["cases_per_gpu"]     # one GPU's share
    pop = ref_arm_pop(w, n_cases)                               # bounded sample: ~2e7 updates per step
    from oracle import binding as ob
    ob.build()
    value, dt = oracle_arm(w, n_cases, pop, args.steps, args.warmup, workers)
    sample = (f"{args.steps} generations at {n_cases} cases with a population subsample of {pop} (cost is linear in pop); "
              "worker threads persist across calls (main_cpu.go spawns goroutines per call: kinder to the CPU than the reference)")
    line = {
        "impl": "reference", "metric": "individual-case semantic updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, w, world, scaling),
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": workers, "kind": "port", "sample": sample,
                         "note": "C restatement of main_cpu.go (oracle); go-gsgp itself cannot be built (no Go toolchain)"},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak (default): the workload's case count per GPU; strong: that many cases in total, split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unfused", action="store_true", help="skip the unfused-pipeline leg (roofline.gs_kernel_unfused stays empty)")
    ap.add_argument("--cpu-baseline-gens", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly one JSON line: whatever libraries print on fd 1 (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the gsgp_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from go_gsgp_b200 import Engine, capi, build
    from go_gsgp_b200.host import HostDriver
    if rank == 0:
        build.build()                                       # no-op when the in-tree .so is current
    if dist is not None:
        dist.barrier()                                      # one builder: the ranks share the file

    w = WORKLOADS[args.workload]
    P, V = w["pop"], w["nvar"]
    scaling = args.scaling or ("strong" if w.get("strong") else "weak")
    if scaling == "strong":                                  # the case count is the job's: every rank holds 1/world of it
        n_global = w["cases_per_gpu"]
        n_local = n_global // world                         # (+1 on the first ranks when it does not divide: gsgp_shard_range)
    else:
        n_local = w["cases_per_gpu"]
        n_global = n_local * world
    nrow, nrow_test = int(0.8 * n_global), n_global - int(0.8 * n_global)
    X, y = synth(w["cfg_id"], n_global, V)

    def fresh_nccl_id():
        """an ncclUniqueId serves one communicator: rank 0 draws a new one per context, broadcast to all"""
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    common = dict(population_size=P, nvar=V, nrow=nrow, nrow_test=nrow_test, max_depth_creation=w["depth"],
                  error_measure=capi.RMSE, p_crossover=w["p_xo"], p_mutation=w["p_mut"], rng_seed=1,
                  device=local_rank, rank=rank, world_size=world)
    host = HostDriver(population_size=P, nvar=V, max_depth_creation=w["depth"], p_crossover=w["p_xo"],
                      p_mutation=w["p_mut"], seed=1)
    sym = host.create_T_F()
    n_seeds = w.get("n_seeds", 0)
    init_trees = host.initialize_population(n_seeds)
    # semantic feeding (config 3): seed k = y + N(0, 0.05 + 0.001 k).  One GPU: drawn independently
    # per seed; several GPUs (every rank needs the global vector, millions of cases x 500 seeds): one N(0, 1) vector drawn
    # once, rotated by 7919 k cases and scaled per seed -- the same marginal distribution without 500 more normal draws
    # per engine set-up.  Throughput does not depend on the values.
    seed_rng = np.random.Generator(np.random.PCG64(20261019 + 100 * w["cfg_id"]))
    base_noise = seed_rng.normal(0.0, 1.0, size=n_global) if (n_seeds and world > 1) else None

    seed_cache = {}

    def seed_vector(k):
        if base_noise is not None:
            return y + (0.05 + 0.001 * k) * np.roll(base_noise, 7919 * k)
        if k not in seed_cache:                                # one GPU: 500 x 1M cases = 4 GB, kept for the three engine set-ups
            seed_cache[k] = y + np.random.Generator(np.random.PCG64(20261019 + 100 * w["cfg_id"] + 1 + k)).normal(0.0, 0.05 + 0.001 * k, size=n_global)
        return seed_cache[k]

    init_ms = []

    def setup(rng_mode, pipeline="fused"):
        os.environ["GSGP_PIPELINE"] = pipeline            # read by gsgp_create
        e = Engine(rng_mode=rng_mode, nccl_id=fresh_nccl_id(), **common)
        e.set_dataset(X, y)
        e.set_constants(sym)
        for k in range(n_seeds):                           # main_cpu.go:1383-1388 (one vector at a time: 500 x 8M cases do not sit in host memory 8 times)
            e.seed_semantic(k, seed_vector(k))
        t0 = time.perf_counter()
        e.eval_initial(n_seeds, init_trees)
        fit, fit_test, best = e.fitness_all()
        init_ms.append((time.perf_counter() - t0) * 1e3)    # the initial population, reported apart from the loop
        return e, fit, best

    K, W = args.steps, args.warmup
    updates_per_step = float(P) * float(n_global)
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- value: device-RNG mode, resident in HBM ------------------------------------------------
    e, fit, best = setup(capi.RNG_DEVICE_PHILOX)
    store = {"buffers": e.store_buffers, "buffer_gb": e.store_buffer_bytes / 1e9, "scratch_gb": e.store_scratch_bytes / 1e9}
    exchange = e.exchange
    stream = torch.cuda.ExternalStream(e.stream)
    e.run_generations(W)
    e.sync()
    launches0 = e.launch_count
    e.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    e.run_generations(K)
    ev1.record(stream)
    e.sync()
    barrier()
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1))
    gpu_launches = e.launch_count - launches0
    prof = {k: e.profile_read(k) for k in (capi.K_INTERP, capi.K_GSOP, capi.K_FINALIZE, capi.K_PLAN)}
    e.profile_enable(False)
    # exact algorithmic bytes of the GS launches need the op mix: count it from the last plan
    plan = e.get_plan()
    computed = (~plan["elite"].astype(bool)) & (plan["op"] != capi.OP_REPR)
    frac_computed = float(np.mean(computed))
    e.close()
    value = updates_per_step * K / (ms_dev * 1e-3)

    # ---- the same generations through the unfused pipeline (interp_kernel -> gsop_kernel): the HBM-bound GS
    # kernel on its own, for the roofline target north_star states for it --------------------------------
    Ku = max(3, min(K, 10))
    prof_u = {capi.K_INTERP: (0, 0.0, 0.0), capi.K_GSOP: (0, 0.0, 0.0)}
    if store["buffers"] == 2 and not args.no_unfused:        # the unfused pipeline needs the cur/new pair
        e, _, _ = setup(capi.RNG_DEVICE_PHILOX, "unfused")
        e.run_generations(W)
        e.sync()
        e.profile_enable(True)
        barrier()
        e.run_generations(Ku)
        e.sync()
        barrier()
        prof_u = {k: e.profile_read(k) for k in (capi.K_INTERP, capi.K_GSOP)}
        e.profile_enable(False)
        e.close()

    # ---- e2e: host-replay mode through gsgp_generation() with host buffers -------------------------
    e, fit, best = setup(capi.RNG_HOST_REPLAY)
    h2d = d2h = 0
    # the host's plan builder is pipelined (gsgph_planner_*, include/gsgp_host.h): the next generation's fitness-
    # independent draws (operators, random trees, tournament candidates) are made on a worker thread while the
    # device computes; winners are picked when the fitness arrives.  Same plans and RNG stream as build_plan.
    # With programs=True the worker threads also compile the trees (gsgp_generation_compiled takes them as they are).
    planner_mode = os.environ.get("GSGP_BENCH_PLANNER", "programs")          # "0": serial gsgph_build_plan, "1": draws only
    # one process per GPU shares the box's cores: the planner's compile threads scale down with the world size
    planner_threads = max(1, min(4, (os.cpu_count() or 4) // (4 * world)))
    # GSGP_BENCH_SHARE_PLAN=1 (N > 1, opt-in until measured on a multi-GPU box): rank 0 alone plans and the other ranks take
    # plan + trees + programs from shared memory (gsgph_channel_*) instead of every rank running a planner
    share = world > 1 and planner_mode == "programs" and os.environ.get("GSGP_BENCH_SHARE_PLAN", "0") == "1"
    channel, seq_next = None, 1
    if share:
        from go_gsgp_b200.host import PlanChannel
        ch_name = "/gsgp_bench_%s" % os.environ.get("MASTER_PORT", "0")
        tree_cap = 4 * (1 << max(w["depth"], 1))
        if rank == 0:
            channel = PlanChannel.create(ch_name, P, 2 * P * tree_cap, 2 * P * (tree_cap // 4 + 1), world - 1)
            planner_threads = max(1, min(4, (os.cpu_count() or 4) // 4))
        barrier()
        if rank != 0:
            channel = PlanChannel.attach(ch_name, rank - 1, P, timeout_ms=60000)
    # first the plain loop a Go host that keeps its own rand.* calls would run: build the plan serially on the host
    # (gsgph_build_plan = main_cpu.go's per-k dispatch), then gsgp_generation() compiles, uploads, runs, reads back
    serial = None
    if not share:
        Ks = max(3, min(K, 15))
        for _ in range(3):
            plan, pool = host.build_plan(fit, best)
            fit, fit_test, best = e.generation(plan, pool)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ks):
            plan, pool = host.build_plan(fit, best)
            fit, fit_test, best = e.generation(plan, pool)
        barrier()
        ms_serial = max_over_ranks((time.perf_counter() - t0) * 1e3)
        serial = {"value": updates_per_step * Ks / (ms_serial * 1e-3), "ms_per_step": ms_serial / Ks, "steps": Ks,
                  "what": "serial host loop: gsgph_build_plan (the reference's per-k RNG dispatch) then gsgp_generation, no overlap of host and device"}
    planner = None if (share and rank != 0) else (
        host.planner(programs=planner_mode == "programs", n_threads=planner_threads) if planner_mode != "0" else None)
    def replay_step(count=False):
        nonlocal fit, best, h2d, d2h, seq_next
        if share:                                            # one generation through the channel, stepped from Python
            if rank == 0:
                plan, pool = planner.plan(fit, best)
                progs = planner.programs()
                channel.publish(seq_next, plan, pool, *progs, timeout_ms=60000)
                planner.prefetch(best)
            else:
                plan, pool, *progs = channel.acquire(seq_next, timeout_ms=60000)
            fit, fit_test, best = e.generation_compiled(plan, pool, *progs)
            if rank != 0:
                channel.release(seq_next)
            seq_next += 1
        else:
            if planner is not None:
                plan, pool = planner.plan(fit, best)
                planner.prefetch(best)
            else:
                plan, pool = host.build_plan(fit, best)
            if planner_mode == "programs":
                fit, fit_test, best = e.generation_compiled(plan, pool, *planner.programs())
            else:
                fit, fit_test, best = e.generation(plan, pool)
        if count:
            # uploaded by the call: plan entries, accumulator programs (8 B per function node; a tree of i ints
            # has (i-1)/4 function nodes), program offset/descriptor tables; read back: 2 fitness vectors + best
            nodes = int(((plan["tree0_len"] - 1) // 4)[plan["tree0_len"] > 0].sum()
                        + ((plan["tree1_len"] - 1) // 4)[plan["tree1_len"] > 0].sum())
            h2d = plan.nbytes + 8 * nodes + 16 * P
            d2h = 16 * P + 4
    # the loop itself runs in the compiled host (gsgph_replay_generations: plan -> gsgp_generation_compiled -> fitness
    # back, K times), like a Go / C++ main would; GSGP_BENCH_E2E_LOOP=python steps it from Python instead
    compiled_loop = (planner is not None or share) and os.environ.get("GSGP_BENCH_E2E_LOOP", "compiled") == "compiled"
    fit_test = np.zeros_like(fit)
    def replay(n):
        nonlocal fit, fit_test, best, seq_next
        if compiled_loop and share:
            fit, fit_test, best, _, _ = channel.replay_generations(planner, e, seq_next, n, fit, fit_test, best)
            seq_next += n
        elif compiled_loop:
            fit, fit_test, best, _, _ = planner.replay_generations(e, n, fit, fit_test, best)
        else:
            for _ in range(n):
                replay_step()
    replay(W)
    barrier()
    t0 = time.perf_counter()
    replay(K)
    barrier()
    t1 = time.perf_counter()
    replay_step(count=True)
    ms_e2e = max_over_ranks((t1 - t0) * 1e3)
    e2e_value = updates_per_step * K / (ms_e2e * 1e-3)
    planner_stats = planner.stats() if planner is not None else None
    e.close()
    if channel is not None:
        barrier()                                            # every reader is done before rank 0 unlinks the segment
        channel.close()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = measured_peak()
    n_gs, ms_gs, _ = prof[capi.K_GSOP]
    n_it, ms_it, _ = prof[capi.K_INTERP]
    local_updates = float(P) * float(n_local)
    frac_xo = float(np.mean(computed & (plan["op"] == capi.OP_XO)))
    frac_mut = float(np.mean(computed & (plan["op"] == capi.OP_MUT)))
    fused = n_it == 0
    if fused:
        # gen_kernel: interpreter + GS operators + error sums in one launch, R stays on chip:
        # XO reads p1, p2 and writes the child (24 B); MUT reads the parent and writes the child (16 B); copy 16 B
        bpu = 24.0 * frac_xo + 16.0 * frac_mut + 16.0 * (1.0 - frac_xo - frac_mut)
        kname = "gen_kernel (tree interpreter + GS crossover/mutation/reproduction + error reduction, fused)"
    else:
        bpu = 32.0 * frac_computed + 16.0 * (1.0 - frac_computed)
        kname = "gsop_kernel (GS crossover/mutation/reproduction + error reduction)"
    achieved = local_updates * bpu * K / (ms_gs * 1e-3) / 1e9 if ms_gs > 0 else None
    # SURVEY 8(d)'s accounting of the same launches: kernel (c) reads the parents AND R (32 B per computed update) -- the
    # fused kernel does that work too, with R served from registers instead of HBM
    bpu_8d = 32.0 * frac_computed + 16.0 * (1.0 - frac_computed)
    achieved_8d = local_updates * bpu_8d * K / (ms_gs * 1e-3) / 1e9 if ms_gs > 0 else None
    # dram__bytes_read.sum + dram__bytes_write.sum per launch: NOT measured in this run -- the figure of the committed
    # `ncu --set full` capture of the same workload and kernel (profiles/README.md), null where there is none
    traffic_tab = {("cfg2", True): (1.888e9, "profiles/r02e_gen_kernel_cfg2_ncu_summary.csv"),
                   ("cfg2", False): (2.895e9, "profiles/r01_gsop_kernel_final_ncu_summary.csv")}
    traffic, traffic_src = traffic_tab.get((args.workload, fused), (None, None)) if world == 1 else (None, None)
    n_gu, ms_gu, _ = prof_u[capi.K_GSOP]
    n_iu, ms_iu, _ = prof_u[capi.K_INTERP]
    bpu_u = 32.0 * frac_computed + 16.0 * (1.0 - frac_computed)
    ach_u = local_updates * bpu_u * Ku / (ms_gu * 1e-3) / 1e9 if ms_gu > 0 else None
    gs_unfused = {"kernel": "gsop_kernel (GS crossover/mutation/reproduction + error reduction), unfused pipeline",
                  "bound": "hbm", "achieved": ach_u, "peak": peak, "unit": "GB/s", "frac": (ach_u / peak) if ach_u else None,
                  "bytes_per_update": bpu_u, "launches": n_gu, "ms_per_launch": ms_gu / max(1, n_gu),
                  "interp_kernel_ms_per_launch": ms_iu / max(1, n_iu),
                  "traffic": traffic_tab[("cfg2", False)][0] if (world == 1 and args.workload == "cfg2") else None,
                  "traffic_source": traffic_tab[("cfg2", False)][1] + " (an earlier capture, not this run)" if (world == 1 and args.workload == "cfg2") else None,
                  "note": "same generations with GSGP_PIPELINE=unfused; not the timed `value` path"}
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": (traffic_src + " (an earlier capture of this workload, not measured in this run)") if traffic_src else None,
                "peak_source": peak_src, "bytes_per_update": bpu,
                "accounting": "bytes the fused kernel has to move: XO 24 B (p1, p2, child), MUT 16 B, copy 16 B per update; R never leaves the SM",
                "survey_8d": {"bytes_per_update": bpu_8d, "achieved": achieved_8d, "frac": (achieved_8d / peak) if achieved_8d else None,
                              "accounting": "SURVEY 8(d) kernel (c): 32 B per crossover / mutation update (parents + R read, child written), 16 B per copy"},
                "launches": n_gs, "ms_total": ms_gs, "ms_per_launch": ms_gs / max(1, n_gs), "pipeline": "fused" if fused else "unfused",
                "interp_kernel": {"launches": n_it, "ms_total": ms_it},
                "share_of_step": {"gs": ms_gs / ms_dev if ms_dev else None, "interp": ms_it / ms_dev if ms_dev else None},
                "limiter": ("instruction issue at 7 warps per scheduler (ncu, profiles/r02e_gen_kernel_cfg2_ncu_summary.csv): issue slots 63 %, "
                            "l1tex data pipe 74 % (shared-memory operand fetches of the interpreter), fp64 pipe 49 %, DRAM 39 %; "
                            "see gs_kernel_unfused for the HBM-bound GS kernel") if fused else "HBM",
                "gs_kernel_unfused": gs_unfused}

    # interpreter work per generation (quoted for config 4): random trees and tree nodes evaluated over every
    # case, counted from the last generation's plan (a pointer-prefix array of L ints holds (L + 1) / 2 nodes)
    try:
        lens = np.concatenate([plan["tree0_len"][computed],
                               plan["tree1_len"][computed & (plan["op"] == capi.OP_MUT)]]).astype(np.int64)
        lens = lens[lens > 0]
        n_trees, n_nodes = int(len(lens)), int(np.sum((lens + 1) // 2))
        per_s = float(n_global) * K / (ms_dev * 1e-3)
        roofline["interpreter_work"] = {"trees_per_generation": n_trees, "nodes_per_generation": n_nodes,
                                        "tree_cases_per_s": n_trees * per_s, "node_cases_per_s": n_nodes * per_s,
                                        "note": "whole job, inside the timed generations (fused into gen_kernel)"}
    except Exception as ex:                                  # reporting only: never costs the bench line
        roofline["interpreter_work"] = {"error": repr(ex)}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:                # rank 0's host cores, at one GPU's share of the cases
        from oracle import binding as ob
        ob.build()
        cores = os.cpu_count() or 1
        pop_s = max(8, min(P, int(4e7 // n_local)))
        v, dt = oracle_arm(w, n_local, pop_s, args.cpu_baseline_gens, 0, cores)
        cpu_baseline = {"value": v, "unit": "updates/s", "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_baseline_gens} generations at {n_local} cases, population subsample {pop_s} "
                                  f"(cost linear in pop), {dt:.1f} s; C restatement of main_cpu.go, not go-gsgp"}
        try:                                                 # and with -workers 1 (SURVEY 8d asks for both)
            pop_1 = max(8, pop_s // 8)
            v1, dt1 = oracle_arm(w, n_local, pop_1, args.cpu_baseline_gens, 0, 1)
            cpu_baseline["single_thread"] = {"value": v1, "cores": 1, "sample": f"population subsample {pop_1}, {dt1:.1f} s"}
        except Exception as ex:
            cpu_baseline["single_thread"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "individual-case semantic updates/sec", "value": value, "unit": "updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, w, world, scaling),
            "engine": {"store": store, "exchange": exchange, "parallelism": f"case-sharded x{world}" if world > 1 else "single GPU",
                       "cases_per_gpu": n_local},
            "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / K, "timing": "host wall clock around K generations of the host-replay loop (plan building + gsgp_generation%s with host buffers), loop in %s" % ("_compiled" if planner_mode == "programs" else "", "the compiled host (gsgph_replay_generations)" if compiled_loop else "Python"),
                    "plan_builder": ("pipelined (gsgph_planner_*%s): elite-index guesses right/wrong = %d/%d" %
                                     (", trees compiled ahead: gsgp_generation_compiled" if planner_mode == "programs" else "",
                                      planner_stats["hits"], planner_stats["misses"])) if planner_stats else "serial (gsgph_build_plan)",
                    "plan_sharing": "rank 0 plans, the other ranks read the plan from shared memory (gsgph_channel_*)" if share
                                    else ("every rank builds the same plan" if world > 1 else "n/a"),
                    "serial_host_loop": serial},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "init": {"ms": init_ms[0] if init_ms else None,
                     "what": "evaluate() of the initial population (main_cpu.go:780-795: %d trees over every case + fitness + best), "
                             "host wall clock with the tree upload; not part of `value`" % (P - n_seeds)},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
