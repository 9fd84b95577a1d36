"""Host-side mirror of the reference crate's public API for the hot path.

Reference (argmin-rs/modcholesky v0.2.0)            here
---------------------------------------------------------------------------------------------
struct Decomposition<L,E,P>{l,e,p}  lib.rs:109-119    Decomposition(l, e, p)
trait ModCholeskyGMW81              gmw81.rs:24-32    ModCholeskyGMW81 (mixin; default raises)
trait ModCholeskySE90               se90.rs:24-32     ModCholeskySE90
trait ModCholeskySE99               se99.rs:21-29     ModCholeskySE99
trait GershgorinCircles             gershgorin.rs:15  GershgorinCircles
impl ... for ndarray::Array2<f64>   gmw81.rs:34-41..  Array2 (numpy-backed), plus the new
                                                      batched Array3 entry point.

Every method goes through the C-ABI (include/modchol_b200.h) into the sm_100a kernels.
numpy arrays are passed as HOST buffers (the library stages them through the GPU); torch CUDA
tensors are passed as DEVICE buffers on torch's current stream (no copies, asynchronous).
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any, List, Tuple

import numpy as np

from . import _lib


@dataclass
class Decomposition:
    """lib.rs:109-119: `l` lower triangular (strict upper explicitly zero), `e` the diagonal
    of E in ORIGINAL index order, `p[i]` the original index at permuted position i, such that
    P (A + diag(e)) P^T = L L^T with P[i, p[i]] = 1.  Carries `phase2_start` (index at which
    SE90/SE99 left phase 1, -1 if never) as an extra diagnostic."""
    l: Any
    e: Any
    p: Any
    phase2_start: Any = None

    @staticmethod
    def new(l, e, p):  # lib.rs:115-118
        return Decomposition(l, e, p)


class NotSquareError(AssertionError, ValueError):
    """The reference asserts `self.is_square()` (se90.rs:41, se99.rs:38, gmw81.rs:43)."""


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _dev_tensor(t, name: str, dtype, shape=None, device=None):
    """A torch tensor about to cross the C-ABI as a DEVICE pointer: it must be a CUDA tensor of
    exactly `dtype` (the kernels reinterpret the bytes), on `device` when given, of `shape`
    when given (None entries are free).  Returns it contiguous."""
    import torch
    if not _is_torch(t):
        raise TypeError(f"{name}: expected a torch CUDA tensor next to the other device buffers, "
                        f"got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name}: torch tensors must live on a CUDA device (pass numpy for host "
                         "data); modcholesky_b200 has no CPU path")
    dtypes = dtype if isinstance(dtype, tuple) else (dtype,)
    if t.dtype not in dtypes:
        raise TypeError(f"{name}: expected {' or '.join(str(d) for d in dtypes)}, got {t.dtype}")
    if device is not None and t.device != device:
        raise ValueError(f"{name}: lives on {t.device}, the other buffers on {device}")
    if shape is not None:
        if t.ndim != len(shape) or any(w is not None and int(g) != int(w) for g, w in zip(t.shape, shape)):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.contiguous()


def _factor_numpy(algo: int, mode: int, a: np.ndarray, ngpus: int = 0) -> Decomposition:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b, n = a.shape[0], a.shape[1]
    l = np.empty_like(a)
    e = np.empty((b, n), dtype=np.float64)
    p = np.empty((b, n), dtype=np.uint64)
    k = np.empty((b,), dtype=np.int32)
    L = _lib.lib()
    if ngpus and ngpus != 1:
        rc = L.modchol_factor_batched_multi_gpu_f64(
            algo, mode, b, n, a.ctypes.data, l.ctypes.data, e.ctypes.data, p.ctypes.data,
            k.ctypes.data, int(ngpus) if ngpus > 0 else 0)
    else:
        rc = L.modchol_factor_batched_f64(
            algo, mode, b, n, a.ctypes.data, l.ctypes.data, e.ctypes.data, p.ctypes.data,
            k.ctypes.data, _lib.MEM_HOST, None)
    _lib.check(rc)
    return Decomposition(l, e, p, k)


def _factor_torch(algo: int, mode: int, a) -> Decomposition:
    import torch
    a = _dev_tensor(a, "a", torch.float64)
    b, n = a.shape[0], a.shape[1]
    with torch.cuda.device(a.device):
        l = torch.empty_like(a)
        e = torch.empty((b, n), dtype=torch.float64, device=a.device)
        p = torch.empty((b, n), dtype=torch.int64, device=a.device)  # same bits as u64 (< 2^63)
        k = torch.empty((b,), dtype=torch.int32, device=a.device)
        st = torch.cuda.current_stream(a.device).cuda_stream
        rc = _lib.lib().modchol_factor_batched_f64(
            algo, mode, b, n, a.data_ptr(), l.data_ptr(), e.data_ptr(), p.data_ptr(), k.data_ptr(),
            _lib.MEM_DEVICE, ctypes.c_void_p(st))
    _lib.check(rc)
    return Decomposition(l, e, p, k)


def factor_batched(algo, a, mode="strict", ngpus: int = 0) -> Decomposition:
    """`a`: (batch, n, n) float64, numpy (host) or torch CUDA tensor (device).  Returns a
    Decomposition of (batch,n,n), (batch,n), (batch,n) arrays of the same kind."""
    if a.ndim != 3 or a.shape[1] != a.shape[2]:
        raise NotSquareError(f"expected (batch, n, n), got {tuple(a.shape)}")
    if a.shape[1] < 1:
        raise NotSquareError("n must be >= 1 (the reference panics for n == 0, se99.rs:190)")
    ai, mi = _lib.algo_id(algo), _lib.mode_id(mode)
    if _is_torch(a):
        return _factor_torch(ai, mi, a)
    return _factor_numpy(ai, mi, np.asarray(a), ngpus)


def factor(algo, a, mode="strict") -> Decomposition:
    """One (n, n) matrix: the batch = 1 case of the same kernels."""
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise NotSquareError(f"matrix must be square, got {tuple(a.shape)}")
    d = factor_batched(algo, a[None], mode)
    return Decomposition(d.l[0], d.e[0], d.p[0], d.phase2_start[0])


# ---- the reference's free-function spelling ---------------------------------------------
def mod_cholesky_gmw81(a, mode="strict") -> Decomposition:
    """gmw81.rs:39 -- Gill, Murray & Wright (1981) on one (n,n) or a (batch,n,n) array."""
    return factor("gmw81", a, mode) if a.ndim == 2 else factor_batched("gmw81", a, mode)


def mod_cholesky_se90(a, mode="strict") -> Decomposition:
    """se90.rs:38 -- Schnabel & Eskow (1990)."""
    return factor("se90", a, mode) if a.ndim == 2 else factor_batched("se90", a, mode)


def mod_cholesky_se99(a, mode="strict") -> Decomposition:
    """se99.rs:35 -- Schnabel & Eskow (1999)."""
    return factor("se99", a, mode) if a.ndim == 2 else factor_batched("se99", a, mode)


def gershgorin_circles_batched(a):
    """gershgorin.rs:20-41 for a (batch, n, n) array -> (batch, n, 2) [centre, radius]."""
    if a.ndim != 3 or a.shape[1] != a.shape[2]:
        raise NotSquareError(f"expected (batch, n, n), got {tuple(a.shape)}")
    L = _lib.lib()
    b, n = a.shape[0], a.shape[1]
    if _is_torch(a):
        import torch
        a = _dev_tensor(a, "a", torch.float64)
        with torch.cuda.device(a.device):
            out = torch.empty((b, n, 2), dtype=torch.float64, device=a.device)
            st = torch.cuda.current_stream(a.device).cuda_stream
            rc = L.modchol_gershgorin_circles_batched_f64(b, n, a.data_ptr(), out.data_ptr(),
                                                          _lib.MEM_DEVICE, ctypes.c_void_p(st))
        _lib.check(rc)
        return out
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty((b, n, 2), dtype=np.float64)
    _lib.check(L.modchol_gershgorin_circles_batched_f64(b, n, a.ctypes.data, out.ctypes.data,
                                                        _lib.MEM_HOST, None))
    return out


def verify_batched(a, d: Decomposition):
    """The identity every reference test asserts (se99.rs:221-231), on the GPU, for a whole
    batch: returns (resid, flags); see modchol_verify_batched_f64 in the header."""
    L = _lib.lib()
    b, n = a.shape[0], a.shape[1]
    if _is_torch(a):
        import torch
        a = _dev_tensor(a, "a", torch.float64, (b, n, n))
        dl = _dev_tensor(d.l, "d.l", torch.float64, (b, n, n), a.device)
        de = _dev_tensor(d.e, "d.e", torch.float64, (b, n), a.device)
        dp = _dev_tensor(d.p, "d.p", (torch.int64, torch.uint64), (b, n), a.device)
        with torch.cuda.device(a.device):
            resid = torch.empty((b,), dtype=torch.float64, device=a.device)
            flags = torch.empty((b,), dtype=torch.int32, device=a.device)
            st = torch.cuda.current_stream(a.device).cuda_stream
            rc = L.modchol_verify_batched_f64(
                b, n, a.data_ptr(), dl.data_ptr(), de.data_ptr(), dp.data_ptr(), resid.data_ptr(),
                flags.data_ptr(), _lib.MEM_DEVICE, ctypes.c_void_p(st))
        _lib.check(rc)
        return resid, flags
    a = np.ascontiguousarray(a, dtype=np.float64)
    l = np.ascontiguousarray(d.l, dtype=np.float64)
    e = np.ascontiguousarray(d.e, dtype=np.float64)
    p = np.ascontiguousarray(d.p, dtype=np.uint64)
    resid = np.empty((b,), dtype=np.float64)
    flags = np.empty((b,), dtype=np.int32)
    _lib.check(L.modchol_verify_batched_f64(b, n, a.ctypes.data, l.ctypes.data, e.ctypes.data,
                                            p.ctypes.data, resid.ctypes.data, flags.ctypes.data,
                                            _lib.MEM_HOST, None))
    return resid, flags


def solve_batched(d: Decomposition, b):
    """Solve (A + E) x = b from a batched Decomposition (the Newton step of the reference's
    callers): `b` is (batch, n) or (batch, nrhs, n); returns x of the same shape and kind."""
    L = _lib.lib()
    squeeze = b.ndim == 2
    if squeeze:
        b = b[:, None, :]
    bt, nrhs, n = b.shape
    if _is_torch(b):
        import torch
        b = _dev_tensor(b, "b", torch.float64)
        dl = _dev_tensor(d.l, "d.l", torch.float64, (bt, n, n), b.device)
        dp = _dev_tensor(d.p, "d.p", (torch.int64, torch.uint64), (bt, n), b.device)
        with torch.cuda.device(b.device):
            x = torch.empty_like(b)
            st = torch.cuda.current_stream(b.device).cuda_stream
            rc = L.modchol_solve_batched_f64(bt, n, nrhs, dl.data_ptr(), dp.data_ptr(), b.data_ptr(),
                                             x.data_ptr(), _lib.MEM_DEVICE, ctypes.c_void_p(st))
        _lib.check(rc)
    else:
        b = np.ascontiguousarray(b, dtype=np.float64)
        l = np.ascontiguousarray(d.l, dtype=np.float64)
        p = np.ascontiguousarray(d.p, dtype=np.uint64)
        if l.shape != (bt, n, n) or p.shape != (bt, n):
            raise ValueError(f"b of shape {tuple(b.shape)} does not match the factors "
                             f"{tuple(l.shape)}, {tuple(p.shape)}")
        x = np.empty_like(b)
        _lib.check(L.modchol_solve_batched_f64(bt, n, nrhs, l.ctypes.data, p.ctypes.data,
                                               b.ctypes.data, x.ctypes.data, _lib.MEM_HOST, None))
    return x[:, 0, :] if squeeze else x


def factor_solve_batched(algo, a, b, mode="strict", return_factors: bool = False):
    """Factorise every a[i] with `algo` and solve (A + E) x = b in the same call (for n <= 64 in
    the same kernel: the factor stays in shared memory).  `a`: (batch, n, n); `b`: (batch, n) or
    (batch, nrhs, n); returns x shaped like b, plus the Decomposition when `return_factors`."""
    if a.ndim != 3 or a.shape[1] != a.shape[2]:
        raise NotSquareError(f"expected (batch, n, n), got {tuple(a.shape)}")
    L = _lib.lib()
    ai, mi = _lib.algo_id(algo), _lib.mode_id(mode)
    squeeze = b.ndim == 2
    if squeeze:
        b = b[:, None, :]
    bt, nrhs, n = b.shape
    if bt != a.shape[0] or n != a.shape[1]:
        raise ValueError(f"b of shape {tuple(b.shape)} does not match a of shape {tuple(a.shape)}")
    d = None
    if _is_torch(a):
        import torch
        a = _dev_tensor(a, "a", torch.float64)
        b = _dev_tensor(b, "b", torch.float64, (bt, nrhs, n), a.device)
        with torch.cuda.device(a.device):
            x = torch.empty_like(b)
            ptrs = [None, None, None, None]
            if return_factors:
                l = torch.empty_like(a)
                e = torch.empty((bt, n), dtype=torch.float64, device=a.device)
                p = torch.empty((bt, n), dtype=torch.int64, device=a.device)
                k = torch.empty((bt,), dtype=torch.int32, device=a.device)
                ptrs = [t.data_ptr() for t in (l, e, p, k)]
                d = Decomposition(l, e, p, k)
            st = torch.cuda.current_stream(a.device).cuda_stream
            rc = L.modchol_factor_solve_batched_f64(ai, mi, bt, n, nrhs, a.data_ptr(), b.data_ptr(),
                                                    x.data_ptr(), *ptrs, _lib.MEM_DEVICE,
                                                    ctypes.c_void_p(st))
        _lib.check(rc)
    else:
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        ptrs = [None, None, None, None]
        if return_factors:
            l = np.empty_like(a)
            e = np.empty((bt, n), dtype=np.float64)
            p = np.empty((bt, n), dtype=np.uint64)
            k = np.empty((bt,), dtype=np.int32)
            ptrs = [t.ctypes.data for t in (l, e, p, k)]
            d = Decomposition(l, e, p, k)
        _lib.check(L.modchol_factor_solve_batched_f64(ai, mi, bt, n, nrhs, a.ctypes.data, b.ctypes.data,
                                                      x.ctypes.data, *ptrs, _lib.MEM_HOST, None))
    x = x[:, 0, :] if squeeze else x
    return (x, d) if return_factors else x


def assemble_qdqt_batched(v, d):
    """The reference's bench inputs (benches/bench.rs:48-53) assembled on the GPU:
    A_b = Q_b diag(d_b) Q_b^T with Q_b the product of the Householder reflectors of the vectors
    v[b, 0], v[b, 1], ... (utils.rs:116-121).  `v`: (batch, nrefl, n), `d`: (batch, n); numpy
    (host) or torch CUDA tensors (device).  Returns (batch, n, n), exactly symmetric."""
    L = _lib.lib()
    if v.ndim != 3 or d.ndim != 2 or v.shape[0] != d.shape[0] or v.shape[2] != d.shape[1]:
        raise ValueError(f"expected v (batch, nrefl, n) and d (batch, n), got {tuple(v.shape)}, {tuple(d.shape)}")
    b, m, n = v.shape
    if _is_torch(d):
        import torch
        d = _dev_tensor(d, "d", torch.float64)
        v = _dev_tensor(v, "v", torch.float64, (b, m, n), d.device)
        with torch.cuda.device(d.device):
            a = torch.empty((b, n, n), dtype=torch.float64, device=d.device)
            st = torch.cuda.current_stream(d.device).cuda_stream
            rc = L.modchol_assemble_qdqt_batched_f64(b, n, m, v.data_ptr(), d.data_ptr(), a.data_ptr(),
                                                     _lib.MEM_DEVICE, ctypes.c_void_p(st))
        _lib.check(rc)
        return a
    v = np.ascontiguousarray(v, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    a = np.empty((b, n, n), dtype=np.float64)
    _lib.check(L.modchol_assemble_qdqt_batched_f64(b, n, m, v.ctypes.data, d.ctypes.data, a.ctypes.data,
                                                   _lib.MEM_HOST, None))
    return a


def fp64_peak_tflops():
    """(DFMA, DMMA) issue peaks of the current device in TFLOP/s, measured by the library."""
    x, y = ctypes.c_double(0.0), ctypes.c_double(0.0)
    _lib.check(_lib.lib().modchol_fp64_peak_tflops(ctypes.byref(x), ctypes.byref(y)))
    return float(x.value), float(y.value)


# ---- the reference's trait spelling -----------------------------------------------------
class ModCholeskyGMW81:
    """gmw81.rs:24-32: the default body panics with "Not implemented!"."""

    def mod_cholesky_gmw81(self) -> Decomposition:
        raise NotImplementedError("Not implemented!")


class ModCholeskySE90:
    """se90.rs:24-32."""

    def mod_cholesky_se90(self) -> Decomposition:
        raise NotImplementedError("Not implemented!")


class ModCholeskySE99:
    """se99.rs:21-29."""

    def mod_cholesky_se99(self) -> Decomposition:
        raise NotImplementedError("Not implemented!")


class GershgorinCircles:
    """gershgorin.rs:15-18."""

    def gershgorin_circles(self) -> List[Tuple[float, float]]:
        raise NotImplementedError("Not implemented!")


def set_default_mode(mode: str) -> str:
    """Arithmetic mode of the trait methods (`Array2(...).mod_cholesky_se99()` takes no arguments,
    like the reference's): "strict" (bit-identical to the reference's CPU arithmetic, the initial
    default; also MODCHOL_DEFAULT_MODE=strict) or "fast" (the north star's tolerances: p equal off
    rounding ties, E to 1e-10, reconstruction to 1e-12; 2.5x the throughput at n = 32).  Returns
    the previous setting -- the same switch as `modcholesky::set_mode` of the Rust facade."""
    _lib.mode_id(mode)   # validates
    prev = _ArrayBase.mode
    _ArrayBase.mode = mode
    return prev


class _ArrayBase:
    mode = "fast" if __import__("os").environ.get("MODCHOL_DEFAULT_MODE", "").lower() == "fast" else "strict"

    def __init__(self, data, mode: str | None = None):
        self.data = data if _is_torch(data) else np.asarray(data, dtype=np.float64)
        if mode is not None:
            self.mode = mode

    def is_square(self) -> bool:
        return self.data.shape[-1] == self.data.shape[-2]


class Array2(_ArrayBase, ModCholeskyGMW81, ModCholeskySE90, ModCholeskySE99, GershgorinCircles):
    """Stand-in for `ndarray::Array2<f64>` carrying the reference's trait impls
    (gmw81.rs:34-41, se90.rs:34-40, se99.rs:31-37, gershgorin.rs:20)."""

    def __init__(self, data, mode=None):
        super().__init__(data, mode)
        if self.data.ndim != 2:
            raise ValueError("Array2 needs a 2-D array")

    def _go(self, algo):
        if not self.is_square():
            raise NotSquareError("assertion failed: self.is_square()")
        return factor(algo, self.data, self.mode)

    def mod_cholesky_gmw81(self) -> Decomposition:
        return self._go("gmw81")

    def mod_cholesky_se90(self) -> Decomposition:
        return self._go("se90")

    def mod_cholesky_se99(self) -> Decomposition:
        return self._go("se99")

    def gershgorin_circles(self):
        if not self.is_square():
            raise NotSquareError("assertion failed: self.is_square()")
        out = gershgorin_circles_batched(self.data[None])[0]
        if _is_torch(out):
            out = out.cpu().numpy()
        return [(float(c), float(r)) for c, r in out]


class Array3(_ArrayBase):
    """The new batched entry point: `Array3<f64>` of shape (batch, n, n) ->
    Decomposition<Array3<f64>, Array2<f64>, Array2<usize>>."""

    def __init__(self, data, mode=None):
        super().__init__(data, mode)
        if self.data.ndim != 3:
            raise ValueError("Array3 needs a 3-D array")

    def _go(self, algo, ngpus=0):
        if not self.is_square():
            raise NotSquareError("assertion failed: self.is_square()")
        return factor_batched(algo, self.data, self.mode, ngpus)

    def mod_cholesky_gmw81_batched(self, ngpus=0) -> Decomposition:
        return self._go("gmw81", ngpus)

    def mod_cholesky_se90_batched(self, ngpus=0) -> Decomposition:
        return self._go("se90", ngpus)

    def mod_cholesky_se99_batched(self, ngpus=0) -> Decomposition:
        return self._go("se99", ngpus)

    def gershgorin_circles_batched(self):
        return gershgorin_circles_batched(self.data)
