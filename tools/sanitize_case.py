"""Small run through every kernel family (for compute-sanitizer memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import modcholesky_b200 as m

rng = np.random.default_rng(0)
for fam in ("reg", "wsm"):
    m.set_small_kernel_family(fam)
    for algo in ("gmw81", "se90", "se99"):
        for n in (1, 2, 7, 16, 23, 32, 33, 48, 64, 70, 170):
            for mode in ("strict", "fast"):
                b = 9 if n <= 64 else 3
                a = rng.uniform(-1, 1, size=(b, n, n))
                a = a + np.transpose(a, (0, 2, 1))
                d = m.factor_batched(algo, torch.from_numpy(a).cuda(), mode=mode)
                r, f = m.verify_batched(torch.from_numpy(a).cuda(), d)
                torch.cuda.synchronize()
                ok = (f == 0) & (r < 1e-12)
                if algo == "se90":   # reference quirk: sqrt of a negative pivot gives NaN (se90.rs:84)
                    ok = ok | ((f & 8) != 0)
                assert bool(ok.all()), (fam, algo, n, mode)
a = rng.uniform(-1, 1, size=(300, 12, 12)); a = a + np.transpose(a, (0, 2, 1))
m.factor_batched("se99", a)                      # host-pointer pipeline
m.gershgorin_circles_batched(a)
print("sanitize_case ok")
