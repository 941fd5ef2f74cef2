"""gsgph_channel_*: one rank plans, the other ranks of the node take plan + tree pool + programs from shared memory
(the host broadcasts the same plan to every GPU).  CPU only: real processes, real shared memory; the
device calls of the shared replay loop are answered by tests/stub_device.c."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.test_host_driver import _fitness_walk

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "channel_worker.py")


@pytest.fixture(scope="module", autouse=True)
def _needs_posix_shared_memory():
    from go_gsgp_b200.host import PlanChannel
    try:
        PlanChannel.create(f"/gsgp_test_probe_{os.getpid()}", 1, 1, 1, 0).close()
    except OSError:
        pytest.skip("no POSIX shared memory in this environment (shm_open failed)")


def _name(tag):
    return f"/gsgp_test_{tag}_{os.getpid()}"


def test_channel_carries_every_plan_to_every_reader():
    from go_gsgp_b200.host import HostDriver, PlanChannel
    from tests.channel_worker import digest
    P, G = 120, 40
    name = _name("plans")
    ch = PlanChannel.create(name, P, 2 * P * 260, 2 * P * 70, 2)
    readers = [subprocess.Popen([sys.executable, WORKER, "reader", name, str(i), str(P), str(G), str(i)], stdout=subprocess.PIPE, text=True)
               for i in range(2)]                              # reader 1 is slow: the writer must wait before reusing a slot
    host = HostDriver(population_size=P, nvar=6, seed=11)
    host.create_T_F(); host.initialize_population(0)
    pl = host.planner(programs=True)
    want = []
    for seq, (fit, best) in enumerate(_fitness_walk(P, G, 5, 0.4), start=1):
        plan, pool = pl.plan(fit, best)
        prog, off, desc = pl.programs()
        ch.publish(seq, plan, pool, prog, off, desc, timeout_ms=20000)
        want.append(digest((plan, pool, prog, off, desc)))
        pl.prefetch(best)
    for r in readers:
        out, _ = r.communicate(timeout=60)
        assert r.returncode == 0
        assert json.loads(out) == want
    ch.close(); host.close()
    assert not os.path.exists("/dev/shm" + name)               # the writer unlinks the segment


def test_channel_errors_and_timeouts():
    from go_gsgp_b200.host import PlanChannel, PLAN_DTYPE
    with pytest.raises(OSError):
        PlanChannel.attach(_name("missing"), 0, 4, timeout_ms=50)
    name = _name("err")
    ch = PlanChannel.create(name, 4, 16, 8, 1)
    with pytest.raises(OSError):
        PlanChannel.attach(name, 1, 4, timeout_ms=50)          # only reader 0 exists
    rd = PlanChannel.attach(name, 0, 4, timeout_ms=1000)
    plan = np.zeros(4, PLAN_DTYPE)
    z = np.zeros(8, np.int32)
    with pytest.raises(RuntimeError, match=r"\(3\)"):
        ch.publish(1, plan, np.zeros(17, np.int32), np.zeros(0, np.uint32), z, z)      # pool larger than the segment
    with pytest.raises(RuntimeError, match=r"\(5\)"):
        rd.acquire(1, timeout_ms=30)                                                  # nothing published yet
    ch.publish(1, plan, np.arange(5, dtype=np.int32), np.arange(6, dtype=np.uint32), z, z + 1)
    ch.publish(2, plan, np.arange(3, dtype=np.int32), np.zeros(0, np.uint32), z, z)
    with pytest.raises(RuntimeError, match=r"\(5\)"):
        ch.publish(3, plan, z[:1], np.zeros(0, np.uint32), z, z, timeout_ms=30)        # slot of 1 not released yet
    p, pool, prog, off, desc = rd.acquire(1)
    assert pool.tolist() == [0, 1, 2, 3, 4] and prog.tolist() == [0, 1, 2, 3, 4, 5] and desc.tolist() == [1] * 8
    rd.release(1)
    ch.publish(3, plan, z[:1], np.zeros(0, np.uint32), z, z, timeout_ms=1000)
    with pytest.raises(RuntimeError, match=r"\(6\)"):
        rd.acquire(1, timeout_ms=30)                                                  # overwritten by 3
    assert rd.acquire(2)[1].tolist() == [0, 1, 2]
    rd.close(); ch.close()


def test_shared_replay_loop_across_processes(tmp_path):
    """gsgph_replay_generations_shared: a writer and two readers, each a process with its own (stub) device context,
    run 30 generations; every process's generation calls received the same plans, trees and programs as a single
    process running gsgph_replay_generations, and end on the same fitness."""
    stub = tmp_path / "libstub_device.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(stub), os.path.join(HERE, "stub_device.c")])
    name = _name("replay")
    env = dict(os.environ, LD_PRELOAD=":".join(x for x in (os.environ.get("LD_PRELOAD"), str(stub)) if x), N_READERS="2")
    P, G = 60, 30
    procs = {role: subprocess.Popen([sys.executable, WORKER, "replay", str(stub), role, name, str(P), str(G)],
                                    stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for role in ("writer", "reader0", "reader1", "solo")}
    res = {}
    for role, p in procs.items():
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, (role, err[-2000:])
        res[role] = json.loads(out.strip().splitlines()[-1])
        assert res[role]["rc"] == 0, role
    for role in ("writer", "reader0", "reader1"):
        assert res[role]["hash"] == res["solo"]["hash"], role                     # identical plans / trees / programs, every generation
        assert res[role]["hist"] == res["solo"]["hist"] and res[role]["fit"] == res["solo"]["fit"] and res[role]["best"] == res["solo"]["best"]
