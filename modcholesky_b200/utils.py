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


def householder_vector(dim: int, seed: int):
    """The vector w of utils.rs:118: `dim` samples of U[-1, 1] from XorShift([seed; 16])."""
    rng = XorShiftRng([seed] * 16)
    dist = UniformInclusiveF64(-1.0, 1.0)
    return np.array([dist.sample(rng) for _ in range(dim)], dtype=np.float64)


def random_householder(dim: int, seed: int):
    """utils.rs:116-121: I - 2 w w^T / ||w||^2 with w ~ U[-1, 1] from XorShift([seed;16])."""
    w = householder_vector(dim, seed).reshape(dim, 1)
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


def bench_input_factors(dim: int, householder_seeds=(2, 9, 90), diag_seed: int = 23,
                        eig_range=(-1.0, 1000.0), min_neg_eigenvalues: int = 1):
    """The ingredients of the reference's bench matrices (benches/bench.rs:48-53): the
    Householder vectors of q1, q2, q3 as a (3, dim) array and the diagonal of d as (dim,).
    `modcholesky_b200.assemble_qdqt_batched` builds A = Q D Q^T from them on the GPU;
    `bench_input_dense` is the reference's own dense construction."""
    v = np.stack([householder_vector(dim, s) for s in householder_seeds])
    d = np.diag(random_diagonal(dim, eig_range, min_neg_eigenvalues, diag_seed)).copy()
    return v, d


def bench_input_dense(dim: int, householder_seeds=(2, 9, 90), diag_seed: int = 23,
                      eig_range=(-1.0, 1000.0), min_neg_eigenvalues: int = 1):
    """benches/bench.rs:48-53 literally: tmp = q1.dot(q2.dot(q3)); a = tmp.dot(d.dot(tmp.t()))."""
    qs = [random_householder(dim, s) for s in householder_seeds]
    tmp = qs[-1]
    for q in reversed(qs[:-1]):
        tmp = q @ tmp
    d = random_diagonal(dim, eig_range, min_neg_eigenvalues, diag_seed)
    return tmp @ (d @ tmp.T)


class _XorShiftLanes:
    """`XorShiftRng` for all 256 one-byte seeds at once (one numpy lane per seed)."""

    def __init__(self):
        s = np.arange(256, dtype=np.uint64)
        w = (s | (s << np.uint64(8)) | (s << np.uint64(16)) | (s << np.uint64(24))).astype(np.uint32)
        w[0] = 0xBAD5EED   # the all-zero seed is replaced (rand_xorshift 0.3 from_seed)
        self.x, self.y, self.z, self.w = w.copy(), w.copy(), w.copy(), w.copy()

    def next_u32(self):
        t = self.x ^ (self.x << np.uint32(11))
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = self.w ^ (self.w >> np.uint32(19)) ^ (t ^ (t >> np.uint32(8)))
        return self.w

    def next_u64(self):
        lo = self.next_u32().astype(np.uint64)
        hi = self.next_u32().astype(np.uint64)
        return (hi << np.uint64(32)) | lo


def _uniform_lanes(rng: _XorShiftLanes, dist: UniformInclusiveF64):
    bits = (rng.next_u64() >> np.uint64(12)) | np.uint64(1023 << 52)
    return (bits.view(np.float64) - 1.0) * dist.scale + dist.low


def householder_vector_table(dim: int):
    """householder_vector(dim, seed) for every seed 0..255 as a (256, dim) array."""
    rng, dist = _XorShiftLanes(), UniformInclusiveF64(-1.0, 1.0)
    return np.stack([_uniform_lanes(rng, dist) for _ in range(dim)], axis=1)


def diagonal_table(dim: int, eig_range=(-1.0, 1000.0), min_neg_eigenvalues: int = 1):
    """diag(random_diagonal(dim, eig_range, min_neg_eigenvalues, seed)) for every seed 0..255 as
    a (256, dim) array (utils.rs:126-147)."""
    rng, dist = _XorShiftLanes(), UniformInclusiveF64(*eig_range)
    w = np.stack([_uniform_lanes(rng, dist) for _ in range(dim)], axis=1)
    idxs = np.tile(np.arange(dim), (256, 1))
    left = dim
    rows = np.arange(256)
    for _ in range(min_neg_eigenvalues):
        f = (rng.next_u64() >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        idxidx = np.floor(f * float(left - 1)).astype(np.int64)
        idx = idxs[rows, idxidx]
        for r in range(256):                      # idxs.remove(idxidx)
            idxs[r, idxidx[r]:left - 1] = idxs[r, idxidx[r] + 1:left]
        left -= 1
        g = (rng.next_u64() >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        w[rows, idx] = g - 1.0
    return w


def bench_input_batch(dim: int, batch: int, first: int = 0):
    """Factors of `batch` distinct matrices of the reference's bench family: matrix t uses the
    Householder seeds (2 + t, 9 + 3 t, 90 + 7 t) mod 256 and the diagonal seed (23 + 11 t) mod
    256, so t = 0 IS the reference's bench matrix (benches/bench.rs:48-53).  Returns
    v (batch, 3, dim) and d (batch, dim) for `assemble_qdqt_batched`."""
    hv, dg = householder_vector_table(dim), diagonal_table(dim)
    t = np.arange(first, first + batch)
    v = np.stack([hv[(2 + t) % 256], hv[(9 + 3 * t) % 256], hv[(90 + 7 * t) % 256]], axis=1)
    return np.ascontiguousarray(v), np.ascontiguousarray(dg[(23 + 11 * t) % 256])
