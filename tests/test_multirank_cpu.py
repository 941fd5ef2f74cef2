"""World-size-2 (gloo, CPU) test of the multi-GPU protocol's host-side logic: the case sharding of
gsgp_shard_range, the per-rank partial error sums, the all-gather + fixed rank-order sum that makes
fitness bit-identical on every rank, and replicated selection staying in lock-step.  The CUDA path
does the same with ncclAllGather (csrc/gsgp_b200.cu launch_finalize); there is no CPU product path,
so the per-shard arithmetic here is numpy over the oracle's semantics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob
from tests.util import synth_dataset


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from go_gsgp_b200 import Engine
        X, y, nrow = synth_dataset(1, 1250, 5, nrow=1000)
        N = len(y)
        cfg = ob.make_config(population_size=40, rng_seed=3, error_measure=ob.RMSE)
        o = ob.Oracle(cfg, X, y, nrow)                      # replicated host: same seed on every rank
        o.create_T_F(); o.initialize_population(0); o.evaluate()
        tr_lo, tr_hi = Engine.shard_range(nrow, rank, world)
        te_lo, te_hi = Engine.shard_range(N - nrow, rank, world)
        fits = []
        for _ in range(3):
            o.generation()
            sem = o.semantics()                                            # [P][N] new generation
            # this rank's partial squared-error sums (train, test) for every individual
            part = np.stack([((y[tr_lo:tr_hi] - sem[:, tr_lo:tr_hi]) ** 2).sum(axis=1),
                             ((y[nrow + te_lo:nrow + te_hi] - sem[:, nrow + te_lo:nrow + te_hi]) ** 2).sum(axis=1)])
            gathered = [torch.zeros(part.shape, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(part))
            tot = np.zeros_like(part)
            for r in range(world):                                         # fixed rank order
                tot = tot + gathered[r].numpy()
            fit = np.sqrt(tot[0] / nrow)
            fit_test = np.sqrt(tot[1] / (N - nrow))
            ref, ref_t = o.fit(), o.fit_test()
            plan = o.last_plan()
            computed = (plan["elite"] == 0) & (plan["op"] != ob.OP_REPR)
            assert np.allclose(fit[computed], ref[computed], rtol=1e-12, atol=0)
            assert np.allclose(fit_test[computed], ref_t[computed], rtol=1e-12, atol=0)
            fits.append(fit.tobytes())
        o.close()
        q.put((rank, (tr_lo, tr_hi, te_lo, te_hi), fits))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_fitness_is_identical_on_every_rank():
    ob.build()
    from go_gsgp_b200 import build
    build.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, f0), (_, s1, f1) = res
    assert s0[0] == 0 and s0[1] == s1[0] and s1[1] == 1000            # train rows partitioned, contiguous
    assert s0[2] == 0 and s0[3] == s1[2] and s1[3] == 250             # test rows likewise
    assert f0 == f1                                                    # bit-identical fitness on both ranks
