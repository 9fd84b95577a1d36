"""ctypes loader for oracle/libmodchol_oracle.so -- TEST INFRASTRUCTURE ONLY.

The shared object is the plain-C restatement of the reference crate's arithmetic
(oracle/modchol_oracle.c, which cites the reference file:line per function).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmodchol_oracle.so")
ALGOS = {"gmw81": 0, "se90": 1, "se99": 2}
_lib = None


def build_oracle(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "modchol_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build_oracle()
    lib = ctypes.CDLL(_SO)
    dp = ctypes.POINTER(ctypes.c_double)
    up = ctypes.POINTER(ctypes.c_uint64)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.oracle_modchol.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, up, ip, dp, dp]
    lib.oracle_modchol.restype = ctypes.c_int
    lib.oracle_modchol_batched.argtypes = [
        ctypes.c_int, ctypes.c_int64, ctypes.c_int, dp, dp, dp, up, ip, dp, ctypes.c_int]
    lib.oracle_modchol_batched.restype = ctypes.c_int
    lib.oracle_gershgorin_circles.argtypes = [ctypes.c_int, dp, dp]
    lib.oracle_gershgorin_circles.restype = None
    lib.oracle_eigenvalues_2x2.argtypes = [ctypes.c_double] * 4 + [dp, dp]
    lib.oracle_eigenvalues_2x2.restype = None
    lib.oracle_max_threads.restype = ctypes.c_int
    lib.oracle_tau.restype = ctypes.c_double
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@dataclass
class OracleResult:
    l: np.ndarray          # (batch, n, n) or (n, n)
    e: np.ndarray          # (batch, n) or (n,)   original index order
    p: np.ndarray          # (batch, n) or (n,)   uint64
    phase2_start: np.ndarray  # int32, -1 = never left phase 1 / GMW81
    margin: np.ndarray     # smallest relative decision gap (see modchol_oracle.c)
    threads: int = 1


def _algo_id(algo) -> int:
    return ALGOS[algo] if isinstance(algo, str) else int(algo)


def factor(algo, a) -> OracleResult:
    """One matrix through the oracle."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == a.shape[1], "matrix must be square"
    r = factor_batched(algo, a[None], threads=1)
    return OracleResult(r.l[0], r.e[0], r.p[0], r.phase2_start[0], r.margin[0], 1)


def factor_batched(algo, a, threads: int = 0) -> OracleResult:
    """(batch, n, n) through the oracle looped over the batch on `threads` host threads
    (0 = all online cores)."""
    lib = _load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 3 and a.shape[1] == a.shape[2], "expected (batch, n, n)"
    b, n = a.shape[0], a.shape[1]
    l = np.empty_like(a)
    e = np.empty((b, n), dtype=np.float64)
    p = np.empty((b, n), dtype=np.uint64)
    k = np.empty(b, dtype=np.int32)
    mg = np.empty(b, dtype=np.float64)
    rc = lib.oracle_modchol_batched(
        _algo_id(algo), b, n, _dp(a), _dp(l), _dp(e),
        p.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
        k.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _dp(mg), int(threads))
    if rc < 0:
        raise RuntimeError(f"oracle_modchol_batched failed: {rc}")
    return OracleResult(l, e, p, k, mg, rc)


def gershgorin_circles(a) -> np.ndarray:
    lib = _load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    out = np.empty((n, 2), dtype=np.float64)
    lib.oracle_gershgorin_circles(n, _dp(a), _dp(out))
    return out


def eigenvalues_2x2(m):
    lib = _load()
    hi, lo = ctypes.c_double(), ctypes.c_double()
    lib.oracle_eigenvalues_2x2(float(m[0][0]), float(m[0][1]), float(m[1][0]), float(m[1][1]),
                               ctypes.byref(hi), ctypes.byref(lo))
    return hi.value, lo.value


def max_threads() -> int:
    return int(_load().oracle_max_threads())


def tau() -> float:
    return float(_load().oracle_tau())
