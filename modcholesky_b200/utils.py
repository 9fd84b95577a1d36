"""Host utilities mirroring `modcholesky::utils` (utils.rs).  These are O(n^2) host-side
helpers the reference's tests use around the hot path; they are numpy, not CUDA (SURVEY.md
section 2 rows 5, 7, 8)."""
from __future__ import annotations

import math

import numpy as np


def eigenvalues_2x2(mat):
    """utils.rs:14-27 -> (hi, lo)."""
    a, b, c, d = float(mat[0][0]), float(mat[0][1]), float(mat[1][0]), float(mat[1][1])
    h = -(a + d) / 2.0
    t = math.sqrt(h * h - a * d + b * c) if (h * h - a * d + b * c) >= 0 else float("nan")
    l1 = (a + d) / 2.0 + t
    l2 = (a + d) / 2.0 - t
    return (l1, l2) if l1 > l2 else (l2, l1)


def swap_columns(mat, idx1, idx2):
    """utils.rs:30-38 (in place)."""
    mat[:, [idx1, idx2]] = mat[:, [idx2, idx1]]


def swap_rows(mat, idx1, idx2):
    """utils.rs:41-49 (in place)."""
    mat[[idx1, idx2], :] = mat[[idx2, idx1], :]


def index_of_largest(c) -> int:
    """utils.rs:52-70: first strict maximum; NaN never replaces the running maximum."""
    mx, mi = c[0], 0
    for i, ci in enumerate(c):
        if ci > mx:
            mx, mi = ci, i
    return mi


def index_of_largest_abs(c) -> int:
    """utils.rs:73-91."""
    mx, mi = abs(c[0]), 0
    for i, ci in enumerate(c):
        if abs(ci) > mx:
            mx, mi = abs(ci), i
    return mi


def index_to_permutation_mat(idxs):
    """utils.rs:94-101: P[i, idxs[i]] = 1."""
    n = len(idxs)
    m = np.zeros((n, n))
    for i in range(n):
        m[i, int(idxs[i])] = 1.0
    return m


def diag_mat_from_arr(arr):
    """utils.rs:104-113."""
    return np.diag(np.asarray(arr))


# ---- bit-exact restatement of the reference's seeded generators (utils.rs:116-147) -------
# rand_xorshift 0.3 XorShiftRng + rand 0.8 Uniform<f64>::new_inclusive / Standard f64.
_M32 = 0xFFFFFFFF


class XorShiftRng:
    """rand_xorshift 0.3.0 `XorShiftRng::from_seed([seed; 16])`."""

    def __init__(self, seed_bytes):
        assert len(seed_bytes) == 16
        w = [int.from_bytes(bytes(seed_bytes[4 * i:4 * i + 4]), "little") for i in range(4)]
        if all(v == 0 for v in w):
            w = [0xBAD_5EED, 0xBAD_5EED, 0xBAD_5EED, 0xBAD_5EED]
        self.x, self.y, self.z, self.w = w

    def next_u32(self) -> int:
        t = (self.x ^ (self.x << 11)) & _M32
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & _M32
        return self.w

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo

    def gen_f64(self) -> float:
        """rand 0.8 `Standard` for f64: 53 random bits scaled by 2^-53."""
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))


class UniformInclusiveF64:
    """rand 0.8.x `Uniform::<f64>::new_inclusive(low, high)` (UniformFloat)."""

    def __init__(self, low: float, high: float):
        max_rand = 1.0 - np.finfo(np.float64).eps       # largest value of the [1,2)-1 sample
        scale = (high - low) / max_rand
        while scale * max_rand + low > high:
            scale = float(np.nextafter(scale, -np.inf))
        self.low, self.scale = low, scale

    def sample(self, rng: XorShiftRng) -> float:
        bits = (rng.next_u64() >> 12) | (1023 << 52)
        value1_2 = np.frombuffer(np.uint64(bits).tobytes(), dtype=np.float64)[0]
        return float((value1_2 - 1.0) * self.scale + self.low)


def random_householder(dim: int, seed: int):
    """utils.rs:116-121: I - 2 w w^T / ||w||^2 with w ~ U[-1, 1] from XorShift([seed;16])."""
    rng = XorShiftRng([seed] * 16)
    dist = UniformInclusiveF64(-1.0, 1.0)
    w = np.array([dist.sample(rng) for _ in range(dim)], dtype=np.float64).reshape(dim, 1)
    denom = 0.0
    for x in w[:, 0]:
        denom = denom + x * x
    return np.eye(dim) - (2.0 / denom) * (w @ w.T)


def random_diagonal(dim: int, eig_range, min_neg_eigenvalues: int, seed: int):
    """utils.rs:126-147."""
    lo, hi = eig_range
    rng = XorShiftRng([seed] * 16)
    dist = UniformInclusiveF64(lo, hi)
    w = np.array([dist.sample(rng) for _ in range(dim)], dtype=np.float64)
    idxs = list(range(dim))
    for _ in range(min_neg_eigenvalues):
        idxidx = int(math.floor(rng.gen_f64() * float(len(idxs) - 1)))
        idx = idxs.pop(idxidx)
        w[idx] = rng.gen_f64() - 1.0
    return np.diag(w)
