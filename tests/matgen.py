"""Seeded test-matrix families (numpy, host).  Families follow SURVEY.md section 8(d):
F1 symmetric-uniform (SE90/SE99 enter phase 2 at k=0), F2 bench-style Q D Q^T
(benches/bench.rs:46-57 shape: three Householder reflections, eigenvalues in (-1, 1000), one
negative), F3 shifted Wishart (phase switch near n/2), PD (never leaves phase 1), plus
tie-heavy small-integer matrices that stress the lowest-index tie-break."""
import numpy as np


def sym_uniform(rng, b, n):
    a = rng.uniform(-1, 1, size=(b, n, n))
    return np.tril(a) + np.transpose(np.tril(a, -1), (0, 2, 1))


def shifted_wishart(rng, b, n, shift=0.05):
    w = rng.standard_normal((b, n, n))
    a = w @ np.transpose(w, (0, 2, 1)) / n - shift * np.eye(n)
    return (a + np.transpose(a, (0, 2, 1))) / 2


def pos_def(rng, b, n):
    w = rng.standard_normal((b, n, n))
    a = w @ np.transpose(w, (0, 2, 1)) + n * np.eye(n)
    return (a + np.transpose(a, (0, 2, 1))) / 2


def bench_style(rng, b, n):
    out = np.empty((b, n, n))
    for i in range(b):
        q = np.eye(n)
        for _ in range(3):
            w = rng.uniform(-1, 1, size=(n, 1))
            q = q @ (np.eye(n) - 2.0 / (w.T @ w).item() * (w @ w.T))
        d = rng.uniform(-1, 1000, size=n)
        d[rng.integers(n)] = rng.uniform(-1, 0)
        a = q @ np.diag(d) @ q.T
        out[i] = (a + a.T) / 2
    return out


def tie_heavy(rng, b, n):
    a = rng.integers(-3, 4, size=(b, n, n)).astype(np.float64)
    a = np.tril(a) + np.transpose(np.tril(a, -1), (0, 2, 1))
    # duplicate a row/column pair in half of the matrices (exact pivot ties, like A6)
    if n >= 3:
        for i in range(0, b, 2):
            r, s = n - 1, n - 2
            a[i, r, :] = a[i, s, :]
            a[i, :, r] = a[i, :, s]
            a[i, r, r] = a[i, s, s]
            a[i, r, s] = a[i, s, r] = a[i, s, s]
    return a


FAMILIES = {
    "uniform": sym_uniform,
    "wishart": shifted_wishart,
    "pd": pos_def,
    "bench": bench_style,
    "ties": tie_heavy,
}


def make(family, seed, b, n):
    return np.ascontiguousarray(FAMILIES[family](np.random.default_rng(seed), b, n))
