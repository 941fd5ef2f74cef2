"""Where does a device-mode generation's time go besides gen_kernel?  (run on the GPU box)"""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from go_gsgp_b200 import Engine, capi
from go_gsgp_b200.host import HostDriver
from bench import synth
P, V, N = 1000, 20, 100000
X, y = synth(2, N, V); nrow = 80000
host = HostDriver(population_size=P, nvar=V, seed=1); sym = host.create_T_F(); trees = host.initialize_population(0)
def mk():
    e = Engine(population_size=P, nvar=V, nrow=nrow, nrow_test=N - nrow, error_measure=capi.RMSE, rng_mode=capi.RNG_DEVICE_PHILOX)
    e.set_dataset(X, y); e.set_constants(sym); e.eval_initial(0, trees); e.fitness_all()
    return e
for prof in (False, True):
    e = mk()
    st = torch.cuda.ExternalStream(e.stream)
    e.run_generations(10); e.sync()
    e.profile_enable(prof)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    a.record(st); t0 = time.perf_counter(); e.run_generations(K); t1 = time.perf_counter(); b.record(st); e.sync()
    ms = a.elapsed_time(b) / K
    line = f"profiling={prof}: {ms:.4f} ms/generation (host enqueue {1e3*(t1-t0)/K:.4f} ms/generation)"
    if prof:
        for name, k in (("gs", capi.K_GSOP), ("finalize", capi.K_FINALIZE), ("plan", capi.K_PLAN)):
            n, tot, _ = e.profile_read(k); line += f"  {name}: {tot/K:.4f} ms ({n//K} launches)"
    print(line)
    e.close()
