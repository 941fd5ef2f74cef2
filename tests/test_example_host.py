"""examples/gsgp_b200_main.cpp: a compiled C++ host over the C-ABI (what main_b200.go does from Go).  CPU: it builds,
links against libgsgp_b200.so and refuses to run without a device.  GPU: its output files equal, line for line, what the
Python mirror produces for the same seed (same host RNG, same plans, same kernels)."""
import os
import subprocess

import numpy as np
import pytest

from go_gsgp_b200 import build
from tests.util import synth_dataset


def write_dataset(path, X, y):
    with open(path, "w") as f:
        f.write(f"{X.shape[1]}\n{X.shape[0]}\n")
        for r, t in zip(X, y):
            f.write("\t".join(repr(float(v)) for v in r) + "\t" + repr(float(t)) + "\n")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ds")
    X, y, nrow = synth_dataset(4, 900, 5, nrow=700)
    write_dataset(d / "train.txt", X[:nrow], y[:nrow])
    write_dataset(d / "test.txt", X[nrow:], y[nrow:])
    return d, X, y, nrow


def test_example_host_builds_and_needs_a_device(files):
    import torch
    exe = build.build_host()
    assert os.access(exe, os.X_OK)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    d, *_ = files
    out = subprocess.run([exe, str(d / "train.txt"), str(d / "test.txt"), str(d), "1", "10"], capture_output=True, text=True)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr


def _stub_fit(gen, k, which):                                  # tests/stub_device.c: stub_fit
    return ((k * 2654435761 + gen * 40503 + which * 97) % 2 ** 32 % 1000) / 8.0 + 1.0


def test_example_host_host_half_with_a_stub_device(files, tmp_path):
    """The example's HOST half end to end on a box without a GPU: tests/stub_device.c (an LD_PRELOAD test stub, a fixed
    hash instead of GSGP) answers the gsgp_* calls, so the program's own work is what is compared with the Python
    mirror driven by the same fake fitness -- Go's random stream, the pipelined planner (plans equal gsgph_build_plan's,
    visible through the contributions), contributions.txt, the asynchronous fitness / semantic writers."""
    from go_gsgp_b200.host import Contributions, HostDriver, format_float, format_semantic
    d, X, y, nrow = files
    exe = build.build_host()
    here = os.path.dirname(os.path.abspath(__file__))
    stub = tmp_path / "libstub_device.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(stub), os.path.join(here, "stub_device.c")])
    gens, pop, seed = 40, 50, 1
    out = subprocess.run([exe, str(d / "train.txt"), str(d / "test.txt"), str(tmp_path), str(gens), str(pop), str(seed)],
                         capture_output=True, text=True, env=dict(os.environ, LD_PRELOAD=":".join(x for x in (os.environ.get("LD_PRELOAD"), str(stub)) if x)))
    assert out.returncode == 0, out.stderr
    host = HostDriver(population_size=pop, nvar=5, seed=seed)
    host.create_T_F(); host.initialize_population(0)
    contrib = Contributions(pop, 0)
    N = len(y)
    ftr, fte, str_, con = [], [], [], []
    for gen in range(gens + 1):
        fit = np.array([_stub_fit(gen, k, 0) for k in range(pop)])
        best = int(np.argmin(fit))                             # first minimum = best_individual's strict comparison
        if gen:
            contrib.apply(plan)
        con.append(contrib.format(best))
        ftr.append(format_float(fit[best])); fte.append(format_float(_stub_fit(gen, best, 1)))
        str_.append(format_semantic((gen + best * 0.5 + np.arange(N) * 0.25)[:nrow]))    # stub_device.c: gsgp_get_semantic
        plan, _ = host.build_plan(fit, best)
        plan = plan.copy()
    assert (tmp_path / "contributions.txt").read_text().split() == con
    assert (tmp_path / "fitnesstrain.txt").read_text().split() == ftr
    assert (tmp_path / "fitnesstest.txt").read_text().split() == fte
    assert (tmp_path / "semantic_train.txt").read_text().split() == str_
    assert int(con[-1]) > 1                                    # crossovers happened: the best's lineage count grew
    import re
    times = (tmp_path / "executiontime.txt").read_text().split()   # Go Duration strings, one per line, increasing
    assert len(times) == gens + 1 and all(re.fullmatch(r"(\d+h)?(\d+m)?\d+(\.\d+)?(ns|µs|ms|s)", t) for t in times)
    host.close()


@pytest.mark.gpu
@pytest.mark.parametrize("linsc", [0, 1])
def test_example_host_matches_python_mirror(files, linsc):
    from go_gsgp_b200 import Engine, capi
    from go_gsgp_b200.host import Contributions, HostDriver, format_float, format_semantic
    d, X, y, nrow = files
    exe = build.build_host()
    gens, pop, seed = 5, 40, 3
    out = subprocess.run([exe, str(d / "train.txt"), str(d / "test.txt"), str(d), str(gens), str(pop), str(seed), str(linsc)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    host = HostDriver(population_size=pop, nvar=5, seed=seed)
    sym = host.create_T_F()
    trees = host.initialize_population(0)
    e = Engine(population_size=pop, nvar=5, nrow=nrow, nrow_test=len(y) - nrow, error_measure=capi.RMSE,
               use_linear_scaling=bool(linsc))
    e.set_dataset(X, y); e.set_constants(sym); e.eval_initial(0, trees)
    fit, fit_test, best = e.fitness_all()
    contrib = Contributions(pop, 0)
    ftr, fte, str_, ste, con = [], [], [], [], []
    def record():
        con.append(contrib.format(best))
        ftr.append(format_float(fit[best])); fte.append(format_float(fit_test[best]))
        s = e.get_semantic(best)
        str_.append(format_semantic(s[:nrow])); ste.append(format_semantic(s[nrow:]))
    record()
    for _ in range(gens):
        plan, pool = host.build_plan(fit, best)
        fit, fit_test, best = e.generation(plan, pool)
        contrib.apply(plan)
        record()
    e.close()
    assert (d / "fitnesstrain.txt").read_text().split() == ftr
    assert (d / "fitnesstest.txt").read_text().split() == fte
    assert (d / "semantic_train.txt").read_text().split() == str_
    assert (d / "semantic_test.txt").read_text().split() == ste
    assert (d / "contributions.txt").read_text().split() == con and con[0] == "1"   # no seeds: one slot, GP itself
    assert len(ftr) == gens + 1                                # gens+1 lines (main_cpu.go:1404-1413,1494-1505)
