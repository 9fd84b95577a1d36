"""In-tree build of the CUDA extension (explicit nvcc, sm_100a only, objects in parallel)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libmodchol_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",          # no implicit contraction anywhere; FAST kernels fuse explicitly
    "-Xcompiler", "-fPIC",
]

# (source, object name, extra defines)
UNITS = [("modchol_api.cu", "api.o", []), ("bsm_inst.cu", "bsm.o", []), ("tpm_inst.cu", "tpm.o", [])] + [
    ("lwm_inst.cu", f"lwm_a{a}.o", [f"-DMODCHOL_INST_ALGO={a}"]) for a in (0, 1, 2)
] + [
    ("warp_inst.cu", f"warp_a{a}_s{s}.o", [f"-DMODCHOL_INST_ALGO={a}", f"-DMODCHOL_INST_STRICT={s}"])
    for a in (0, 1, 2) for s in (1, 0)
] + [
    ("wsm_inst.cu", f"wsm_a{a}_s{s}.o", [f"-DMODCHOL_INST_ALGO={a}", f"-DMODCHOL_INST_STRICT={s}"])
    for a in (0, 1, 2) for s in (1, 0)
] + [
    ("csm_inst.cu", f"csm_a{a}_s{s}.o", [f"-DMODCHOL_INST_ALGO={a}", f"-DMODCHOL_INST_STRICT={s}"])
    for a in (0, 1, 2) for s in (1, 0)
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                  + [os.path.join(os.path.dirname(_HERE), "include", "modchol_b200.h"),
                     os.path.abspath(__file__)])


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def _env():
    env = dict(os.environ)
    env.pop("CC", None)   # $CC may point at a gcc without the full runtime; nvcc finds g++ itself
    env.pop("CXX", None)
    return env


def build_library(force: bool = False, verbose: bool = False, jobs: int | None = None,
                  variant: str | None = None, extra: list[str] | None = None) -> str:
    """variant: build into csrc/build_<variant>/ and libmodchol_b200.<variant>.so with the extra
    nvcc flags (kernel-tuning experiments and the bounds-checked "dbg" build; selected at run
    time with MODCHOL_LIB)."""
    obj_dir, lib = OBJ, LIB
    if variant:
        obj_dir = os.path.join(CSRC, "build_" + variant)
        lib = os.path.join(CSRC, f"libmodchol_b200.{variant}.so")
        if not force and os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in sources()):
            return lib
    elif not force and not is_stale():
        return lib
    os.makedirs(obj_dir, exist_ok=True)
    exe, env = nvcc(), _env()

    extra = (extra or []) + os.environ.get("MODCHOL_NVCC_EXTRA", "").split()   # e.g. -DMODCHOL_WIP_MINB=5

    def compile_one(unit):
        src, obj, defs = unit
        defs = defs + extra
        cmd = [exe] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", "-o", os.path.join(obj_dir, obj), os.path.join(CSRC, src)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {obj}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=jobs or min(len(UNITS), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(compile_one, UNITS))
    if verbose:
        for obj, log in logs:
            print(f"==== {obj}\n{log}")
    subprocess.check_call([exe, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib]
                          + [os.path.join(obj_dir, u[1]) for u in UNITS], env=env)
    return lib


DEBUG_VARIANT = ("dbg", ["-DMODCHOL_DEBUG_BOUNDS=1"])   # shared-memory accessors bounds-checked (wsm_se.cuh)


def build_debug_library(force: bool = False) -> str:
    return build_library(force=force, variant=DEBUG_VARIANT[0], extra=DEBUG_VARIANT[1])


if __name__ == "__main__":
    import sys
    args = [a for a in sys.argv[1:] if a != "-v"]
    # python -m modcholesky_b200.build [-v] [variant -Dflag ...]
    print(build_library(force=True, verbose="-v" in sys.argv, variant=args[0] if args else None,
                        extra=args[1:]))
