"""The interpreter's program builder (csrc/program.hpp) on the CPU: every reference tree array is
turned into an accumulator program; running that program with numpy must reproduce the oracle's
eval_arrays (main_cpu.go:755-775) bit for bit, including protected division and NaN/Inf flow."""
import numpy as np
import pytest

from oracle import binding as ob
from go_gsgp_b200 import build
from go_gsgp_b200.host import compile_tree
from tests.util import synth_dataset, toy_fixture

A_ADD, A_SUB, A_MUL, A_DIV, A_RSUB, A_RDIV, A_MOV = range(7)
K_LOADA, K_PUSH, K_POP = 8, 16, 32


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()
    ob.build()


def pdiv(num, den):
    with np.errstate(all="ignore"):
        return np.where(den == 0.0, 1.0, num / np.where(den == 0.0, 1.0, den))


def run_program(code, ab, cols):
    """cols: list of operand columns (variables then constants), each an ndarray over cases."""
    acc = np.zeros_like(cols[0])
    stack = {}
    with np.errstate(all="ignore"):
        for c, w in zip(code.tolist(), ab.tolist()):
            level, aop = c >> 11, c & 7                      # program.hpp kLevelShift
            if c & K_PUSH:
                stack[level] = acc.copy()
            if c & K_LOADA:
                acc = cols[w & 0xFFFF].copy()
            t = stack[level] if c & K_POP else cols[w >> 16]
            acc = {A_ADD: lambda: acc + t, A_SUB: lambda: acc - t, A_MUL: lambda: acc * t, A_DIV: lambda: pdiv(acc, t),
                   A_RSUB: lambda: t - acc, A_RDIV: lambda: pdiv(t, acc), A_MOV: lambda: t.copy()}[aop]()
    return acc


def _function_positions(tree):
    """marks the array slots that hold function ids (walks the pointer-prefix form)."""
    mark = np.zeros(len(tree), bool)
    todo = [0]
    while todo:
        p = todo.pop()
        if tree[p] < 4:
            mark[p] = True
            todo += [int(tree[p + 1]), int(tree[p + 2])]
    return mark


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)]) and \
        np.array_equal(np.isnan(a), np.isnan(b))


def test_toy_tree_program():
    # (+ x0 (* x1 x0)) from SURVEY Appendix B: one TT instruction for the product, one T+acc for the sum
    code, ab, need = compile_tree([0, 3, 4, 4, 2, 7, 8, 5, 4], 6)
    assert need == 0 and len(code) == 2
    assert code[0] == (K_LOADA | A_MUL) and ab[0] == (1 | (0 << 16))          # acc = x1 ; t = x0 ; acc *= t
    assert code[1] == A_ADD and ab[1] == (0 << 16)                            # acc = x0 + acc (commutes)
    X, y, nrow = toy_fixture()
    out = run_program(code, ab, [X[:, 0], X[:, 1]])
    assert same_bits(out, X[:, 0] + X[:, 1] * X[:, 0])
    # a lone terminal (zero_depth trees)
    code, ab, need = compile_tree([5], 6)
    assert len(code) == 1 and code[0] == A_MOV and ab[0] >> 16 == 1


@pytest.mark.parametrize("depth,nconst,zero_depth", [(6, 0, 0), (8, 4, 0), (10, 3, 1), (3, 0, 0)])
def test_programs_reproduce_oracle_tree_semantics(depth, nconst, zero_depth):
    X, y, nrow = synth_dataset(1, 257, 5, nrow=200)
    X = X.copy()
    X[::7, 2] = 0.0            # exact zeros: protected division and 0*inf paths
    X[::11, 3] = -0.0
    cfg = ob.make_config(population_size=4, max_depth_creation=depth, num_random_constants=nconst, rng_seed=11,
                         zero_depth=zero_depth)
    o = ob.Oracle(cfg, X, y, nrow)
    o.create_T_F()
    sym = o.symbol_values()
    cols = [X[:, v] for v in range(5)] + [np.full(len(X), sym[4 + 5 + k]) for k in range(nconst)]
    needs = []
    for _ in range(400):
        tree = o.create_grow_tree_arrays()
        code, ab, need = compile_tree(tree, len(sym))
        n_func = int(np.sum(_function_positions(tree)))
        assert len(code) == max(1, n_func)                  # one instruction per function node
        out = run_program(code, ab, cols)
        ref = o.semantic_evaluate_array(tree)
        assert same_bits(out, ref)
        levels = [c >> 11 for c in code.tolist() if c & (K_PUSH | K_POP)]
        assert all(l < need for l in levels)
        needs.append(need)
    o.close()
    # Sethi-Ullman ordering keeps almost every grow tree within two spill levels
    assert np.mean(np.asarray(needs) <= 2) > 0.9


def test_malformed_trees_are_rejected():
    # the last three share nodes between parents (a DAG): they would expand into more instructions than len / 4 + 1
    for bad in ([0, 3, 4, 4], [0, 1, 2], [0, 3, 4, 99, 4], [0, 0, 0, 4, 4], [9], [0, 3, 3, 1, 6, 6, 4], [0, 3, 3, 4],
                [0, 3, 3, 1, 6, 6, 2, 9, 9, 4]):
        with pytest.raises(ValueError):
            compile_tree(bad, 6)


def test_generated_division_ptx_is_current():
    """go_gsgp_b200/csrc/interp_div.inc.h is generated: the checked-in text must be what tools/gen_interp_ptx.py renders."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_interp_ptx", os.path.join(root, "tools", "gen_interp_ptx.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(gen.TARGET).read() == gen.render(gen.DEFAULT_GROUP)
    # gen_kernel's threaded-code interpreter (interp_body.inc.h) comes from the same generator
    assert open(gen.TARGET_BODY).read() == gen.render_body(gen.DEFAULT_GROUP)
