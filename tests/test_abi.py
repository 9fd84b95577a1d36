"""CPU-only checks of the boundary: the C-ABI library loads, exports exactly what
include/modchol_b200.h declares, rejects bad arguments, and refuses to compute without a GPU
(there is no CPU fallback).  Also the host-side mirror of utils.rs."""
import ctypes
import os
import re

import numpy as np
import pytest

import modcholesky_b200 as m
from modcholesky_b200 import _lib, utils
from modcholesky_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    build_library()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "modchol_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(modchol_[a-z0-9_]+)\s*\(", src)))


def test_header_and_library_agree():
    syms = _header_symbols()
    assert syms == sorted(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"


def test_version_and_kernel_classes():
    assert m.lib().modchol_version() == 100
    assert m.kernel_class("gmw81", 256) == "cta_packed" and m.kernel_class("gmw81", 257) == "cta_gmem"
    assert m.kernel_class("se99", 128) == "cta_packed" and m.kernel_class("se90", 236) == "cta_packed"
    assert m.kernel_class("se90", 237) == "cta_packed" and m.kernel_class("gmw81", 128) == "cta_packed"
    assert m.kernel_class("se90", 64) == "warp_smem"
    assert m.kernel_class("se99", 32, "fast") == "warp_smem" and m.kernel_class("gmw81", 64) == "warp_smem"
    assert m.kernel_class("gmw81", 65) == "cta_packed"
    assert m.kernel_class("se99", 32, "strict") == "warp_smem" and m.kernel_class("se99", 16, "strict") == "thread"
    assert m.kernel_class("gmw81", 32) == "warp_smem" and m.kernel_class("gmw81", 3, "fast") == "thread"
    assert m.kernel_class("se99", 4096) == "unsupported"


def test_rust_facade_binds_every_entry_point():
    """rust/src/ffi.rs cannot be compiled here (no rustc): at least every C-ABI function of the
    header must be declared there with the same number of parameters."""
    hdr = open(os.path.join(ROOT, "include", "modchol_b200.h")).read()
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    found = re.findall(r"\b(modchol_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert sorted(n for n, _ in found) == sorted(_lib.EXPORTS)
    for name, cargs in found:
        mm = re.search(r"pub fn %s\s*\(([^)]*)\)" % name, ffi, re.S)
        assert mm, f"{name} is not declared in rust/src/ffi.rs"
        nc = 0 if cargs.strip() in ("", "void") else cargs.count(",") + 1
        rargs = mm.group(1).strip().rstrip(",")
        nr = 0 if not rargs else rargs.count(",") + 1
        assert nc == nr, f"{name}: {nc} parameters in the header, {nr} in ffi.rs"


def test_bad_arguments_are_rejected_before_any_cuda_work():
    L = m.lib()
    a = np.zeros((1, 2, 2))
    args = (a.ctypes.data,) * 4
    assert L.modchol_factor_batched_f64(7, 0, 1, 2, *args, None, 0, None) == -1     # algo
    assert L.modchol_factor_batched_f64(2, 9, 1, 2, *args, None, 0, None) == -1     # mode
    assert L.modchol_factor_batched_f64(2, 0, -1, 2, *args, None, 0, None) == -1    # batch
    assert L.modchol_factor_batched_f64(2, 0, 1, 0, *args, None, 0, None) == -1     # n == 0 panics in the reference
    assert L.modchol_factor_batched_f64(2, 0, 1, 5000, *args, None, 0, None) == -3  # > MAX_N
    assert L.modchol_factor_batched_f64(2, 0, 1, 2, None, None, None, None, None, 0, None) == -1
    assert b"null" in L.modchol_last_error()


def test_fused_factor_solve_rejects_bad_arguments():
    L = m.lib()
    a = np.zeros((1, 2, 2))
    b = np.zeros((1, 1, 2))
    pa, pb = a.ctypes.data, b.ctypes.data
    f = L.modchol_factor_solve_batched_f64
    assert f(2, 0, 1, 2, 0, pa, pb, pb, None, None, None, None, 0, None) == -1     # nrhs < 1
    assert f(2, 0, 1, 2, 1, pa, None, pb, None, None, None, None, 0, None) == -1   # no right-hand side
    assert f(2, 0, 1, 2, 1, pa, pb, None, None, None, None, None, 0, None) == -1   # no x
    assert f(9, 0, 1, 2, 1, pa, pb, pb, None, None, None, None, 0, None) == -1     # algo
    assert f(2, 0, 1, 2, 1, pa, pb, pb, None, None, None, None, 7, None) == -1     # mem_kind
    assert f(2, 0, 1, 5000, 1, pa, pb, pb, None, None, None, None, 0, None) == -3  # > MAX_N
    if m.device_count() == 0:   # every output optional, but still no CPU path
        assert f(2, 0, 1, 2, 1, pa, pb, pb, None, None, None, None, 0, None) == -4
    with pytest.raises(ValueError):
        m.factor_solve_batched("se99", np.zeros((2, 3, 3)), np.zeros((2, 4)))


def test_non_square_raises_like_the_reference_assert():
    with pytest.raises(m.NotSquareError):
        m.Array2(np.zeros((2, 3))).mod_cholesky_se99()
    with pytest.raises(AssertionError):
        m.mod_cholesky_gmw81(np.zeros((2, 3)))


def test_trait_defaults_panic_not_implemented():
    # gmw81.rs:29-31, se90.rs:29-31, se99.rs:26-28
    class Foo(m.ModCholeskyGMW81, m.ModCholeskySE90, m.ModCholeskySE99):
        pass
    for meth in ("mod_cholesky_gmw81", "mod_cholesky_se90", "mod_cholesky_se99"):
        with pytest.raises(NotImplementedError, match="Not implemented!"):
            getattr(Foo(), meth)()


@pytest.mark.skipif(m.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    a = np.array([[1.0, 1.0, 2.0], [1.0, 1.0, 3.0], [2.0, 3.0, 1.0]])
    with pytest.raises(m.ModCholError) as ei:
        m.Array2(a).mod_cholesky_se99()
    assert ei.value.code == -4


def test_product_package_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import modcholesky_b200; "
            "assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules)" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, "modcholesky_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libmodchol_oracle" not in txt, f


# ---- utils.rs mirror (host side), reference tests utils.rs:153-241 -----------------------
def test_utils_swaps_and_argmax(golden):
    g = golden["utils"]
    a = np.array(g["swap"]["a"])
    b = a.copy(); utils.swap_columns(b, 1, 2)
    assert np.array_equal(b, np.array(g["swap"]["swap_columns_1_2"]))
    b = a.copy(); utils.swap_rows(b, 1, 2)
    assert np.array_equal(b, np.array(g["swap"]["swap_rows_1_2"]))
    b = a.copy(); utils.swap_rows(b, 1, 2); utils.swap_columns(b, 1, 2)
    assert np.array_equal(b, np.array(g["swap"]["swap_rows_then_columns_1_2"]))
    t = g["index_of_largest"]
    assert utils.index_of_largest(np.diag(np.array(t["a"]))[t["j"]:]) + t["j"] == t["expect"]
    t = g["index_of_largest_abs"]
    assert utils.index_of_largest_abs(np.diag(np.array(t["a"]))[t["j"]:]) + t["j"] == t["expect"]
    assert utils.index_of_largest(np.array([2.0, 5.0, 5.0, 1.0])) == 1          # first maximum
    assert utils.index_of_largest(np.array([np.nan, 5.0, 7.0])) == 0            # NaN head sticks
    assert utils.eigenvalues_2x2([[2.0, 1.0], [1.0, 2.0]]) == (3.0, 1.0)


def test_seeded_generators_are_bit_exact(golden):
    g = golden["utils"]
    q = utils.random_householder(2, 84)
    assert np.abs(q - np.array(g["random_householder_2_84"]["expect"])).max() < 1e-19
    d = utils.random_diagonal(2, (-1.0, 1.0), 0, 128)
    assert np.abs(d - np.array(g["random_diagonal_2_all_pos_128"]["expect"])).max() < 1e-19
    d = utils.random_diagonal(2, (-1.0, 1.0), 1, 128)
    assert np.abs(d - np.array(g["random_diagonal_2_one_neg_128"]["expect"])).max() < 1e-19
    q = utils.random_householder(12, 2)
    assert np.abs(q @ q.T - np.eye(12)).max() < 1e-14 and np.abs(q - q.T).max() == 0.0
