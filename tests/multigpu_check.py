"""Run under torchrun (one rank per GPU): case-sharded engines in lock-step with the (global) oracle.
Every rank passes the same dataset / plan / trees; each keeps its shard; fitness must come back
bit-identical on every rank and within 1e-9 of the oracle; the local semantic shards must match."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as ob                      # noqa: E402
from go_gsgp_b200 import Engine, capi                 # noqa: E402
from tests.util import synth_dataset, rel_err, same_best         # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_nccl_id():                                   # an ncclUniqueId serves ONE communicator
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    # two shapes, both ragged (shards of unequal size; 8 ranks get 7 tiles each of the first, ~100 of the second)
    shapes = ((20_003, 15_001, 7, 64, 6), (200_011, 160_003, 20, 24, 3))
    for (n_cases, n_train, n_var, P, gens), (rng_kind, mode) in (
            (sh, rm) for sh in shapes for rm in ((ob.RNG_ALFG, capi.RNG_HOST_REPLAY), (ob.RNG_PHILOX, capi.RNG_DEVICE_PHILOX))):
        X, y, nrow = synth_dataset(2, n_cases, n_var, nrow=n_train)
        N, V = X.shape
        cfg = ob.make_config(population_size=P, rng_kind=rng_kind, rng_seed=5, error_measure=ob.RMSE,
                             use_linear_scaling=int(mode == capi.RNG_DEVICE_PHILOX), n_workers=max(1, (os.cpu_count() or 1) // world))
        o = ob.Oracle(cfg, X, y, nrow)
        o.create_T_F()
        e = Engine(population_size=P, nvar=V, nrow=nrow, nrow_test=N - nrow, error_measure=capi.RMSE, rng_seed=5,
                   rng_mode=mode, device=local, rank=rank, world_size=world, nccl_id=fresh_nccl_id(),
                   use_linear_scaling=(mode == capi.RNG_DEVICE_PHILOX),
                   # the device-mode leg also runs on the single-buffer store (what config 5 uses at full size)
                   flags=capi.FLAG_SINGLE_BUFFER_STORE if mode == capi.RNG_DEVICE_PHILOX else 0)
        assert e.store_buffers == (1 if mode == capi.RNG_DEVICE_PHILOX else 2)
        want = os.environ.get("GSGP_EXCHANGE", "p2p")
        assert e.exchange == want, (e.exchange, want)       # the fused peer exchange is the default; NCCL on request
        e.set_dataset(X, y)
        e.set_constants(o.symbol_values())
        o.initialize_population(0)
        e.eval_initial(0, [o.initial_tree(i) for i in range(P)])
        o.evaluate()
        fit, fit_test, best = e.fitness_all()
        assert same_best(best, o) and np.all(rel_err(fit, o.fit()) <= 1e-9)
        tr_lo, tr_hi = Engine.shard_range(nrow, rank, world)
        te_lo, te_hi = Engine.shard_range(N - nrow, rank, world)
        for g in range(gens):
            o.generation()
            if mode == capi.RNG_HOST_REPLAY:
                fit, fit_test, best = e.generation(o.last_plan(), o.last_tree_pool())
            else:
                e.run_generations(1)
                fit, fit_test, best = e.get_fitness()
                plan = e.get_plan()
                ref = o.last_plan()
                for f in ("op", "elite", "p1", "p2"):
                    live = ref["elite"] == 0 if f in ("p1", "p2") else slice(None)
                    assert np.array_equal(plan[f][live], ref[f][live]), (rank, g, f)
            assert same_best(best, o), (rank, g, best, o.index_best)
            assert np.all(rel_err(fit, o.fit()) <= 1e-9) and np.all(rel_err(fit_test, o.fit_test()) <= 1e-9)
            allfit = [torch.zeros(P, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(allfit, torch.from_numpy(fit).cuda())
            for t in allfit:                                              # bit-identical on every rank
                assert torch.equal(t.view(torch.int64), allfit[0].view(torch.int64))
        sem = e.get_semantics()
        ref = o.semantics()
        ref_local = np.concatenate([ref[:, tr_lo:tr_hi], ref[:, nrow + te_lo:nrow + te_hi]], axis=1)
        assert sem.shape == ref_local.shape and np.all(rel_err(sem, ref_local) <= 1e-9)
        e.close(); o.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world} exchange={os.environ.get('GSGP_EXCHANGE', 'p2p')}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
