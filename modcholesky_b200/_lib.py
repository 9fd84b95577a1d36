"""ctypes binding of the C-ABI in include/modchol_b200.h.

The shared library is built IN-TREE (modcholesky_b200/csrc/libmodchol_b200.so) by
modcholesky_b200.build.build_library() / __graft_entry__.build().  There is no CPU or PyTorch
fallback: if the library is missing, or no CUDA device is visible, compute calls raise.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MODCHOL_LIB: another in-tree build of the same sources (kernel-tuning experiments, tools/)
LIB_PATH = os.environ.get("MODCHOL_LIB") or os.path.join(_HERE, "csrc", "libmodchol_b200.so")

GMW81, SE90, SE99 = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
MODE_STRICT, MODE_FAST = 0, 1
ALGO_IDS = {"gmw81": GMW81, "se90": SE90, "se99": SE99}
MODE_IDS = {"strict": MODE_STRICT, "fast": MODE_FAST}

# every symbol include/modchol_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTS = (
    "modchol_factor_batched_f64",
    "modchol_factor_batched_multi_gpu_f64",
    "modchol_gershgorin_circles_batched_f64",
    "modchol_verify_batched_f64",
    "modchol_solve_batched_f64",
    "modchol_factor_solve_batched_f64",
    "modchol_assemble_qdqt_batched_f64",
    "modchol_fp64_peak_tflops",
    "modchol_kernel_class",
    "modchol_set_small_kernel_family",
    "modchol_launch_count",
    "modchol_last_error",
    "modchol_version",
    "modchol_device_count",
    "modchol_shutdown",
)


class ModCholError(RuntimeError):
    """A C-ABI call returned a negative status (the Rust facade turns this into panic!)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"modchol error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "modcholesky_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.modchol_factor_batched_f64.argtypes = [ctypes.c_int, ctypes.c_int, i64, i32, vp, vp, vp, vp, vp,
                                             ctypes.c_int, vp]
    L.modchol_factor_batched_f64.restype = ctypes.c_int
    L.modchol_factor_batched_multi_gpu_f64.argtypes = [ctypes.c_int, ctypes.c_int, i64, i32, vp, vp, vp,
                                                       vp, vp, ctypes.c_int]
    L.modchol_factor_batched_multi_gpu_f64.restype = ctypes.c_int
    L.modchol_gershgorin_circles_batched_f64.argtypes = [i64, i32, vp, vp, ctypes.c_int, vp]
    L.modchol_gershgorin_circles_batched_f64.restype = ctypes.c_int
    L.modchol_verify_batched_f64.argtypes = [i64, i32, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    L.modchol_verify_batched_f64.restype = ctypes.c_int
    L.modchol_solve_batched_f64.argtypes = [i64, i32, i32, vp, vp, vp, vp, ctypes.c_int, vp]
    L.modchol_solve_batched_f64.restype = ctypes.c_int
    L.modchol_factor_solve_batched_f64.argtypes = [ctypes.c_int, ctypes.c_int, i64, i32, i32, vp, vp, vp,
                                                   vp, vp, vp, vp, ctypes.c_int, vp]
    L.modchol_factor_solve_batched_f64.restype = ctypes.c_int
    L.modchol_assemble_qdqt_batched_f64.argtypes = [i64, i32, i32, vp, vp, vp, ctypes.c_int, vp]
    L.modchol_assemble_qdqt_batched_f64.restype = ctypes.c_int
    L.modchol_fp64_peak_tflops.argtypes = [vp, vp]
    L.modchol_fp64_peak_tflops.restype = ctypes.c_int
    L.modchol_kernel_class.argtypes = [ctypes.c_int, ctypes.c_int, i32]
    L.modchol_kernel_class.restype = ctypes.c_char_p
    L.modchol_set_small_kernel_family.argtypes = [ctypes.c_int]
    L.modchol_set_small_kernel_family.restype = ctypes.c_int
    L.modchol_launch_count.restype = i64
    L.modchol_last_error.restype = ctypes.c_char_p
    L.modchol_version.restype = ctypes.c_int
    L.modchol_device_count.restype = ctypes.c_int
    L.modchol_shutdown.restype = None
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise ModCholError(rc, lib().modchol_last_error().decode())


def algo_id(algo) -> int:
    return ALGO_IDS[algo.lower()] if isinstance(algo, str) else int(algo)


def mode_id(mode) -> int:
    return MODE_IDS[mode.lower()] if isinstance(mode, str) else int(mode)
