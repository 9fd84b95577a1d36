"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/modchol_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this.  The product package (modcholesky_b200) never does.
"""
from .oracle import (  # noqa: F401
    ALGOS,
    OracleResult,
    build_oracle,
    eigenvalues_2x2,
    factor,
    factor_batched,
    gershgorin_circles,
    max_threads,
)
