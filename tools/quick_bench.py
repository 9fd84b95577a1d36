"""Quick device-resident throughput probe (not the contract bench; see bench.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modcholesky_b200 as m


def gen(family, batch, n, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if family == "uniform":
        a = torch.rand((batch, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        return torch.tril(a) + torch.tril(a, -1).transpose(1, 2)
    w = torch.randn((batch, n, n), dtype=torch.float64, device="cuda", generator=g)
    a = w @ w.transpose(1, 2)
    if family == "wishart":
        a = a / n - 0.05 * torch.eye(n, dtype=torch.float64, device="cuda")
    else:
        a = a + n * torch.eye(n, dtype=torch.float64, device="cuda")
    return (a + a.transpose(1, 2)) / 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="se99:32:1048576,se99:16:1048576")
    ap.add_argument("--modes", default="strict,fast")
    ap.add_argument("--families", default="uniform,wishart,pd")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    for case in args.cases.split(","):
        algo, n, batch = case.split(":")
        n, batch = int(n), int(batch)
        for fam in args.families.split(","):
            a = gen(fam, batch, n)
            for mode in args.modes.split(","):
                d = m.factor_batched(algo, a, mode=mode)
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record()
                    d = m.factor_batched(algo, a, mode=mode)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t = sorted(ts)[len(ts) // 2]
                rate = batch / (t * 1e-3)
                gbs = rate * (16 * n * n + 16 * n) / 1e9
                k = d.phase2_start
                print(f"{algo} n={n} batch={batch} {fam:8s} {mode:6s} {m.kernel_class(algo, n, mode):9s}"
                      f" {t:9.3f} ms  {rate/1e6:9.3f} M mat/s  {gbs:8.1f} GB/s alg"
                      f"  ({gbs/6551*100:5.1f}% of 6551)  k_mean={k.float().mean().item():.2f}"
                      f" frac_phase2={(k >= 0).float().mean().item():.3f}", flush=True)
            del a, d


if __name__ == "__main__":
    main()
