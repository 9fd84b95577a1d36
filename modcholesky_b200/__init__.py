"""modcholesky_b200 -- B200-native batched modified Cholesky (GMW81 / SE90 / SE99, f64).

A drop-in for the hot path of argmin-rs/modcholesky: same trait names, same
Decomposition {l, e, p}, plus a batched Array3 entry point.  All compute runs in
hand-written sm_100a CUDA kernels behind the C-ABI of include/modchol_b200.h; importing
this package never imports the CPU oracle and there is no CPU fallback.
"""
from . import utils  # noqa: F401
from ._lib import (  # noqa: F401
    GMW81, SE90, SE99, MEM_DEVICE, MEM_HOST, MODE_FAST, MODE_STRICT, ModCholError, lib)
from .api import (  # noqa: F401
    Array2, Array3, Decomposition, assemble_qdqt_batched, fp64_peak_tflops, GershgorinCircles, ModCholeskyGMW81, ModCholeskySE90,
    ModCholeskySE99, NotSquareError, factor, factor_batched, factor_solve_batched,
    gershgorin_circles_batched,
    mod_cholesky_gmw81, mod_cholesky_se90, mod_cholesky_se99, set_default_mode, solve_batched,
    verify_batched)

__version__ = "0.1.0"


def kernel_class(algo, n: int, mode="strict") -> str:
    from . import _lib
    return _lib.lib().modchol_kernel_class(_lib.algo_id(algo), _lib.mode_id(mode), int(n)).decode()


def device_count() -> int:
    return int(lib().modchol_device_count())


def launch_count() -> int:
    return int(lib().modchol_launch_count())


def set_small_kernel_family(family) -> int:
    """0/'auto', 1/'reg' (register-resident), 2/'wsm' (shared-memory-resident), 3/'tpm' (one thread
    per matrix, n <= 8): which kernel family serves n <= 32.  Returns the previous setting."""
    ids = {"auto": 0, "reg": 1, "wsm": 2, "tpm": 3}
    return int(lib().modchol_set_small_kernel_family(ids.get(family, family)))
