"""Randomised parity sweep beyond the test-suite: every n in a range, odd batch sizes, all algorithms
and families; STRICT must be bit-identical to the oracle, FAST within the north-star tolerances.
usage: python tools/stress_parity.py [nmax] [seed]   (needs a GPU; the oracle is the checker)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import matgen
import modcholesky_b200 as m
from oracle import oracle
import test_parity_gpu as T

nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 72
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
sizes = list(range(1, nmax + 1)) + [80, 96, 113, 128, 130, 177, 233, 240, 256, 290, 384, 513]
bad = 0
cases = 0
for n in sizes:
    b = 37 if n <= 72 else 5
    for fam in ("uniform", "wishart", "pd", "bench", "ties"):
        a = matgen.make(fam, seed * 100000 + n, b, n)
        for algo in ("gmw81", "se90", "se99"):
            ref = oracle.factor_batched(algo, a)
            for mode in ("strict", "fast"):
                if mode == "fast" and (fam == "ties" or n < 2):
                    continue
                d = m.factor_batched(algo, torch.from_numpy(a).cuda(), mode=mode)
                got = (d.l.cpu().numpy(), d.e.cpu().numpy(), d.p.cpu().numpy().astype(np.uint64),
                       d.phase2_start.cpu().numpy())
                cases += 1
                try:
                    if mode == "strict":
                        T._assert_bit_exact(f"{algo} n={n} {fam}", got, ref)
                    else:
                        T._assert_tolerance(f"{algo} n={n} {fam}", a, got, ref)
                except AssertionError as e:
                    bad += 1
                    print("FAIL", algo, n, fam, mode, str(e)[:200])
print(f"{cases} cases, {bad} failures")
sys.exit(1 if bad else 0)
