"""Parity (FAST bars of tests/test_parity_gpu.py) and throughput of the large-n kernels.
usage: python tools/check_large.py [algo] [n,n,...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import matgen
import modcholesky_b200 as m
import oracle
from test_parity_gpu import _assert_tolerance, _gpu

oracle.build_oracle()
algo = sys.argv[1] if len(sys.argv) > 1 else "gmw81"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [129, 200, 256, 300, 512, 1000, 1024]
for n in sizes:
    b = 8 if n <= 300 else 3
    for fam in ("uniform", "wishart", "pd"):
        a = matgen.make(fam, 7000 + n, b, n)
        ref = oracle.factor_batched(algo, a)
        got = _gpu(torch, algo, a, "fast")
        try:
            _assert_tolerance(f"{algo} n={n} {fam}", a, got, ref)
            status = "ok"
        except AssertionError as ex:
            status = "FAIL " + str(ex)[:150]
        print(f"{algo} n={n} {fam:8s} {m.kernel_class(algo, n, 'fast'):12s} {status}", flush=True)
    # throughput, device resident
    batch = max(148 * 4, min(4096, (1 << 30) // (8 * n * n)))
    g = torch.Generator(device="cuda").manual_seed(1)
    ad = torch.rand((batch, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    ad = torch.tril(ad) + torch.tril(ad, -1).transpose(1, 2)
    for _ in range(2):
        d = m.factor_batched(algo, ad, mode="fast")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    d = m.factor_batched(algo, ad, mode="fast")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rate = batch / ms * 1e3
    print(f"{algo} n={n} batch={batch}: {ms:9.2f} ms  {rate / 1e6:8.4f} M mat/s  "
          f"{rate * (16 * n * n + 16 * n) / 1e9:7.1f} GB/s alg  {rate * n ** 3 / 3 / 1e12:6.2f} TFLOP/s", flush=True)
    del ad, d
