"""-proto_dump (main_cpu.go:1416-1436,1441-1488,1510-1517): gsgph_proto_* writes pb.Evolution in protobuf wire format.
The files are read back here with python-protobuf through a descriptor restating the reference's evolution.proto, and
compared byte for byte with python-protobuf's own serialization of the same content."""
import numpy as np
import pytest

from oracle import binding as ob
from tests.util import synth_dataset


def _pb():
    """message classes of evolution.proto (package pb): RandomTree, Individual, Population, Evolution"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="evolution.proto", package="pb", syntax="proto3")
    rt = fd.message_type.add(name="RandomTree")                                          # evolution.proto:19-23
    rt.field.add(name="data", number=1, type=F.TYPE_INT32, label=F.LABEL_REPEATED)
    rt.field.add(name="semantic_train", number=2, type=F.TYPE_DOUBLE, label=F.LABEL_REPEATED)
    rt.field.add(name="semantic_test", number=3, type=F.TYPE_DOUBLE, label=F.LABEL_REPEATED)
    ind = fd.message_type.add(name="Individual")                                         # :25-40
    en = ind.enum_type.add(name="Operator")
    for i, n in enumerate(("INIT", "XO", "MUT", "REPR")):
        en.value.add(name=n, number=i)
    ind.field.add(name="op", number=1, type=F.TYPE_ENUM, type_name=".pb.Individual.Operator", label=F.LABEL_OPTIONAL)
    ind.field.add(name="fitness_train", number=2, type=F.TYPE_DOUBLE, label=F.LABEL_OPTIONAL)
    ind.field.add(name="fitness_test", number=3, type=F.TYPE_DOUBLE, label=F.LABEL_OPTIONAL)
    ind.field.add(name="semantic_train", number=4, type=F.TYPE_DOUBLE, label=F.LABEL_REPEATED)
    ind.field.add(name="semantic_test", number=5, type=F.TYPE_DOUBLE, label=F.LABEL_REPEATED)
    ind.field.add(name="contrib", number=6, type=F.TYPE_BYTES, label=F.LABEL_OPTIONAL)
    ind.field.add(name="rt", number=7, type=F.TYPE_MESSAGE, type_name=".pb.RandomTree", label=F.LABEL_REPEATED)
    pop = fd.message_type.add(name="Population")                                         # :42-45
    pop.field.add(name="generation", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    pop.field.add(name="individuals", number=2, type=F.TYPE_MESSAGE, type_name=".pb.Individual", label=F.LABEL_REPEATED)
    evo = fd.message_type.add(name="Evolution")                                          # :47-50
    evo.field.add(name="generations", number=1, type=F.TYPE_MESSAGE, type_name=".pb.Population", label=F.LABEL_REPEATED)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("pb." + n))
            for n in ("RandomTree", "Individual", "Population", "Evolution")}


def test_wire_format_equals_python_protobuf():
    from go_gsgp_b200.host import ProtoDump
    pb = _pb()
    rng = np.random.default_rng(4)
    for nrow, nrow_test in ((5, 3), (4, 0), (0, 0), (300, 77)):
        n = nrow + nrow_test
        d = ProtoDump(nrow, nrow_test)
        want = pb["Evolution"]()
        for gen in (0, 1, 2, 300):
            d.begin_population(gen)
            wp = want.generations.add(generation=gen)
            for k in range(7):
                op = k % 4
                fit = [0.0, -0.0, float("nan"), float("inf"), 1.5e300, 5e-324, -3.25][k]
                fit_test = float(rng.normal())
                sem = rng.normal(size=n)
                if n:
                    sem[0] = 0.0
                contrib = bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8)) if k % 3 else b""
                trees = []
                for t in range({1: 1, 2: 2}.get(op, 0)):
                    if k == 5 and t == 0:
                        trees.append((None, None))                               # rt never assigned: empty RandomTree
                    else:
                        trees.append((rng.integers(-5, 3000, size=int(rng.integers(1, 60))).astype(np.int32), rng.normal(size=n)))
                d.add_individual(op, fit, fit_test, sem, contrib, trees)
                wi = wp.individuals.add(op=op, fitness_train=fit, fitness_test=fit_test, semantic_train=sem[:nrow],
                                        semantic_test=sem[nrow:], contrib=contrib)
                for data, ts in trees:
                    if data is None:
                        wi.rt.add()
                    else:
                        wi.rt.add(data=data.tolist(), semantic_train=ts[:nrow], semantic_test=ts[nrow:])
            d.end_population()
        got = d.tobytes()
        assert got == want.SerializeToString(deterministic=True)
        back = pb["Evolution"].FromString(got)
        assert [p.generation for p in back.generations] == [0, 1, 2, 300]
        assert len(back.generations[3].individuals) == 7 and back.generations[3].individuals[2].op == 2
        d.close()


def test_dump_of_a_run_reads_back(tmp_path):
    """Three generations driven by the oracle's plans: every individual's operator, fitness, semantic and random trees
    (tree array + its semantic) come back from the file; the elite's entry repeats the previous trees (rt1 / rt2 are
    globals in the reference, :217-226, and the elite creates none, :877,918)."""
    from go_gsgp_b200.host import ProtoDump
    pb = _pb()
    X, y, nrow = synth_dataset(1, 90, 5, nrow=60)
    P = 14
    o = ob.Oracle(ob.make_config(population_size=P, rng_seed=3, p_crossover=0.5, p_mutation=0.4), X, y, nrow)
    o.create_T_F(); o.initialize_population(0); o.evaluate()
    o.keep_tree_semantics(True)
    d = ProtoDump(nrow, 90 - nrow)
    d.add_initial_population(o.fit(), o.fit_test(), o.semantics())
    d2 = ProtoDump(nrow, 90 - nrow)                            # the same dump with the trees fetched one by one (device mode)
    d2.add_initial_population(o.fit(), o.fit_test(), o.semantics())
    record = []
    for gen in range(1, 4):
        o.generation()
        plan, pool = o.last_plan().copy(), o.last_tree_pool().copy()
        d.add_generation(gen, plan, pool, o.fit(), o.fit_test(), o.semantics(), o.last_tree_semantic)
        d2.add_generation(gen, plan, None, o.fit(), o.fit_test(), o.semantics(), o.last_tree_semantic,
                          tree=lambda k, w: pool[plan[f"tree{w}_off"][k]: plan[f"tree{w}_off"][k] + plan[f"tree{w}_len"][k]])
        record.append((plan, pool, o.fit().copy(), o.semantics().copy(),
                       {(k, w): o.last_tree_semantic(k, w).copy() for k in range(P) for w in (0, 1)
                        if plan[("tree0_len", "tree1_len")[w]][k] > 0 and not plan["elite"][k]}))
    path = tmp_path / "dump.pb"
    d.write(path)
    assert path.read_bytes() == d.tobytes() == d2.tobytes()
    evo = pb["Evolution"].FromString(path.read_bytes())
    assert [p.generation for p in evo.generations] == [0, 1, 2, 3]
    assert all(i.op == 0 and not i.rt for i in evo.generations[0].individuals)
    last = [None, None]
    seen_elite_with_trees = False
    for (plan, pool, fit, sems, tsem), popm in zip(record, evo.generations[1:]):
        assert len(popm.individuals) == P
        for k, im in enumerate(popm.individuals):
            e = plan[k]
            assert im.op == e["op"] and len(im.rt) == {ob.OP_XO: 1, ob.OP_MUT: 2}.get(int(e["op"]), 0)
            assert (im.fitness_train == fit[k]) or (np.isnan(fit[k]) and np.isnan(im.fitness_train))
            np.testing.assert_array_equal(np.array(im.semantic_train), sems[k][:nrow])
            np.testing.assert_array_equal(np.array(im.semantic_test), sems[k][nrow:])
            if e["op"] in (ob.OP_XO, ob.OP_MUT) and not e["elite"]:
                last[0] = (pool[e["tree0_off"]: e["tree0_off"] + e["tree0_len"]], tsem[(k, 0)])
                if e["op"] == ob.OP_MUT:
                    last[1] = (pool[e["tree1_off"]: e["tree1_off"] + e["tree1_len"]], tsem[(k, 1)])
            elif im.rt:
                seen_elite_with_trees = True
            for w, r in enumerate(im.rt):
                if last[w] is None:
                    assert not r.data and not r.semantic_train
                    continue
                assert list(r.data) == last[w][0].tolist()
                np.testing.assert_array_equal(np.array(r.semantic_train), last[w][1][:nrow])
                np.testing.assert_array_equal(np.array(r.semantic_test), last[w][1][nrow:])
    assert seen_elite_with_trees, "the run should exercise the elite-under-XO/MUT case"
    d.close()


def test_call_order_and_io_errors(tmp_path):
    from go_gsgp_b200.host import ProtoDump
    d = ProtoDump(2, 1)
    with pytest.raises(RuntimeError):
        d.end_population()
    with pytest.raises(RuntimeError):
        d.add_individual(1, 1.0, 2.0, np.zeros(3))
    d.begin_population(0)
    with pytest.raises(RuntimeError):
        d.begin_population(1)
    with pytest.raises(RuntimeError):
        d.add_individual(9, 1.0, 2.0, np.zeros(3))                   # not an Individual_Operator
    with pytest.raises(RuntimeError):
        d.tobytes()                                                  # a population is still open
    d.add_individual(3, 1.0, 2.0, np.array([1.0, 2.0, 3.0]))
    d.end_population()
    assert d.tobytes() == bytes.fromhex("0a32" "1230" "0803" "11000000000000f03f" "190000000000000040"
                                        "2210" "000000000000f03f" "0000000000000040" "2a08" "0000000000000840")
    with pytest.raises(OSError):
        d.write(tmp_path / "no_such_dir" / "x.pb")
    d.close()
