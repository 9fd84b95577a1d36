"""End-to-end throughput of the trait path with pageable numpy buffers (bench.py reports the same
figure as e2e.trait_path_pageable_numpy).  usage: python tools/e2e_pageable.py [batch] [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import modcholesky_b200 as m

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rng = np.random.default_rng(1)
a = rng.uniform(-1, 1, size=(batch, n, n))
a = np.tril(a) + np.transpose(np.tril(a, -1), (0, 2, 1))
arr = m.Array3(a, mode="fast")
arr.mod_cholesky_se99_batched()
for rep in range(3):
    t0 = time.perf_counter()
    d = arr.mod_cholesky_se99_batched()
    dt = time.perf_counter() - t0
    print(f"pageable numpy, fresh outputs: {batch / dt / 1e6:.3f} M matrices/s ({dt * 1e3:.1f} ms)")
