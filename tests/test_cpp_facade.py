"""The C++ host facade (include/modcholesky.hpp): compiles and links against the C-ABI library
on CPU; the reference's doctest runs through it on the GPU box."""
import os
import subprocess

import pytest

from modcholesky_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")


def _build():
    lib = build_library()
    libdir = os.path.dirname(lib)
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(
        ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
         os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), "-o", EXE,
         "-L", libdir, "-lmodchol_b200", f"-Wl,-rpath,{libdir}"], env=env)


def test_facade_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_reference_doctest_through_cpp_facade():
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "facade ok" in r.stdout
