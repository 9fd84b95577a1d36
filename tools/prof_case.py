"""One case, a few launches: the command profiled under ncu (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modcholesky_b200 as m
from quick_bench import gen

algo, n, batch, fam, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
a = gen(fam, batch, n)
for _ in range(reps):
    d = m.factor_batched(algo, a, mode=mode)
    r = m.verify_batched(a, d)   # the on-device acceptance check is part of the profiled command
torch.cuda.synchronize()
print("ok", algo, n, batch, fam, mode, m.kernel_class(algo, n, mode))
