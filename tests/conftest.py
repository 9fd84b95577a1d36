import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_literals.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build_oracle()
    return orc


# ---- the reference's own check helpers (utils.rs:94-113), restated for numpy ----
def perm_matrix(p):
    """utils.rs:94-101 index_to_permutation_mat: P[i, p[i]] = 1."""
    n = len(p)
    m = np.zeros((n, n))
    for i in range(n):
        m[i, int(p[i])] = 1.0
    return m


def identity_residual(a, l, e, p):
    """max |P A P^T + P diag(e) P^T - L L^T|  (the identity every reference test asserts,
    e.g. se99.rs:221-231)."""
    P = perm_matrix(p)
    lhs = P @ a @ P.T + P @ np.diag(e) @ P.T
    return np.abs(lhs - l @ l.T).max()
