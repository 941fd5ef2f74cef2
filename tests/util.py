"""Shared helpers for the parity tests (synthetic data as BASELINE.md section 4 defines it)."""
import numpy as np


def synth_dataset(cfg_id: int, n: int, v: int, train_frac: float = 0.8, nrow: int | None = None):
    """X ~ U[-1,1), y = x0*x1 + 0.5*x2 - x3 + 0.25*x4^2 + N(0,0.01); PCG64(20261019 + cfg)."""
    rng = np.random.Generator(np.random.PCG64(20261019 + cfg_id))
    X = rng.uniform(-1.0, 1.0, size=(n, v))
    cols = [X[:, i] if i < v else np.zeros(n) for i in range(5)]
    y = cols[0] * cols[1] + 0.5 * cols[2] - cols[3] + 0.25 * cols[4] ** 2 + rng.normal(0.0, 0.01, size=n)
    if nrow is None:
        nrow = int(round(train_frac * n))
    return X, y, nrow


def toy_fixture():
    """The 4-row dataset of main_test.go:107-112: 2 train rows, 2 test rows, V=2."""
    X = np.array([[1.1, 2.2], [4.4, 5.5], [1.2, 2.4], [4.1, 2.6]])
    y = np.array([3.3, 9.9, 3.6, 6.7])
    return X, y, 2


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    denom = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(denom > 0, np.abs(a - b) / denom, 0.0)
    e = np.where(both_nan | same_inf, 0.0, e)
    e = np.where(np.isnan(e), np.inf, e)
    return e


def same_best(best, o, tol=1e-11):
    """The CUDA path's best index against the oracle's: equal, or the two candidates' fitness ties to within the
    summation-order noise in the oracle's own values (decisions are exact given equal fitness values)."""
    if best == o.index_best:
        return True
    f = o.fit()
    return bool(rel_err(f[best], f[o.index_best]) <= tol)
