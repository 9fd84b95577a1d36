"""Fused factor+solve versus factor followed by solve (device-resident, CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modcholesky_b200 as m
from quick_bench import gen


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for algo, n, batch in (("se99", 4, 1 << 22), ("se99", 8, 1 << 20), ("se99", 32, 1 << 19), ("se99", 16, 1 << 20), ("se90", 64, 100000), ("gmw81", 32, 1 << 19)):
    a = gen("uniform", batch, n)
    b = torch.randn((batch, n), dtype=torch.float64, device="cuda")
    for mode in ("strict", "fast"):
        t_f = timed(lambda: m.factor_batched(algo, a, mode=mode))
        t_sep = timed(lambda: m.solve_batched(m.factor_batched(algo, a, mode=mode), b))
        t_fused = timed(lambda: m.factor_solve_batched(algo, a, b, mode=mode))
        print(f"{algo} n={n} batch={batch} {mode:6s} factor {batch / t_f / 1e3:8.2f} M/s | factor+solve kernels "
              f"{batch / t_sep / 1e3:8.2f} M/s | fused {batch / t_fused / 1e3:8.2f} M/s")
