"""Arithmetic identities the CUDA kernels rely on, checked in IEEE double on the CPU (numpy follows the same rules).

gen_kernel (kernels.cuh) may evaluate a mutation's two random trees in either order: when tree 1 needs every spill level
and tree 2 does not, the block's setup swaps them and negates the mutation step, i.e. it computes
T + (-ms) * (s2 - s1) where the reference (main_cpu.go:932-936) computes T + ms * (s1 - s2).  The two are the same
double, bit for bit -- s2 - s1 == -(s1 - s2) exactly (rounding to nearest is symmetric) and so is the product's sign --
with one exception that no later arithmetic can see: x - x is +0 in either order, so when the two sigmas are EQUAL the
product's zero changes sign, and if T is itself a zero the child may be -0 where the reference has +0 (or the other way
round).  Semantics are only ever combined linearly and subtracted from the targets: the sign of a zero never reaches a
fitness value.
"""
import numpy as np


def test_mutation_with_swapped_trees_is_the_same_double():
    rng = np.random.default_rng(7)
    n = 200_000
    s1 = 1.0 / (1.0 + np.exp(-rng.normal(0, 30, n)))
    s2 = 1.0 / (1.0 + np.exp(-rng.normal(0, 30, n)))
    # the corners sigma can take: exactly 0.5, exactly 1, 2^-1024 (exp64's MaxFloat64 rule), equal values, tiny differences
    s1[:6] = [0.5, 1.0, 2.0 ** -1024, 0.25, 1.0, np.nextafter(0.5, 1)]
    s2[:6] = [0.5, 2.0 ** -1024, 1.0, 0.25, np.nextafter(1.0, 0), 0.5]
    T = rng.normal(0, 1e3, n) * (rng.random(n) < 0.9)              # some exact zeros, of both signs
    T[:6] = [0.0, 1.0, -0.0, -0.0, 0.0, 0.0]
    ms = rng.uniform(-1, 1, n)
    ms[:3] = [0.0, -0.0, 1.0]
    want = T + ms * (s1 - s2)
    got = T + (-ms) * (s2 - s1)
    zero_of_zero = (s1 == s2) & (T == 0.0)                         # the documented exception: a zero whose sign may differ
    assert zero_of_zero.any() and np.all(want[zero_of_zero] == 0.0) and np.all(got[zero_of_zero] == 0.0)
    assert np.array_equal(want.view(np.int64)[~zero_of_zero], got.view(np.int64)[~zero_of_zero])


def test_sigma_accumulated_from_zero_is_sigma():
    """S = r - 0 (the crossover path of the generation kernels' shared loop) is r itself, the sign of zero included."""
    r = np.array([0.0, -0.0, 1.0, 2.0 ** -1024, 0.5, np.nextafter(1.0, 0)])
    assert np.array_equal((r - 0.0).view(np.int64), r.view(np.int64))
