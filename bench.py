#!/usr/bin/env python
"""bench.py -- the measurement contract of this repo.

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

Metric (BASELINE.json): matrices/s of batched f64 SE99 factorisations at n = 32.
Workload at every N: 2^20 32x32 symmetric indefinite matrices PER GPU (weak scaling; the 8-GPU
point is BASELINE.json configs[4], "8M 32x32 SE99 Hessians sharded across 8xB200").  A step is one
factorisation of the resident batch: read A (8 GiB), write L, e, p (8.25 GiB) -- far larger than
the 126 MB L2, so no flush is needed between steps.

value    : matrices/s with the batch already resident in HBM (CUDA events on the launching
           stream, barrier + synchronize on both sides, max over ranks).
e2e      : the same metric through the public API with HOST (pinned) buffers: the C-ABI
           stages host -> device -> host in overlapped chunks inside the timed region.
roofline : algorithmic bytes (16 n^2 + 16 n per matrix) / kernel time,
           against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
configs  : (N = 1 only) short timed runs of the other BASELINE.json configs -- configs[0] the
           reference's own single-matrix bench shapes (CPU oracle us/call and the GPU batch = 1
           call latency), configs[1] 2^20 x 16^2 SE99, configs[2] 200 000 x 64^2 SE90, configs[3]
           4 096 x 256^2 GMW81 (with the FP64-pipe fraction against the DFMA / DMMA peaks the
           library measures on this device) -- plus the reference's bench family (Q D Q^T,
           12 / 128 / 256, assembled on the device) and the on-device verifier.
cpu_baseline : the CPU oracle (a C restatement of the reference; the Rust crate cannot be
           built here) looped over a bounded sample on all host cores, rank 0, N = 1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MAT = 32
ALGO = "se99"
BATCH_PER_GPU = 1 << 20
E2E_BATCH_PER_GPU = 1 << 18
CPU_SAMPLE = 1 << 18
SEED = 0xC0FFEE
ALG_BYTES = 16 * N_MAT * N_MAT + 16 * N_MAT          # 16 896 B per matrix


def gen_batch(torch, batch, n, seed, device):
    """Family F1: symmetric, entries ~ U[-1, 1] -- indefinite, SE99 enters
    phase 2 at k = 0 (the Gershgorin path)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((batch, n, n), dtype=torch.float64, device=device)
    step = 1 << 17
    for s in range(0, batch, step):
        a = torch.rand((min(step, batch - s), n, n), dtype=torch.float64, device=device, generator=g) * 2 - 1
        out[s:s + a.shape[0]] = torch.tril(a) + torch.tril(a, -1).transpose(1, 2)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, scaled to
    this batch (profiles/r02_traffic.json), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            return t["dram_bytes_per_matrix"] * BATCH_PER_GPU
        except Exception:
            continue
    return None



def kernel_name(m, algo, n, mode):
    """Kernel class reported by the library plus the kernel the dispatcher runs for it."""
    cls = m.kernel_class(algo, n, mode)
    A = algo.upper()
    if mode == "fast" and cls == "warp_smem" and n > 4:
        if n <= 8:
            return f"{cls}: wipp_kernel<{A}, G=8> (implicit pivoting, four matrices per warp, L in tensor memory)"
        if n <= 16:
            return f"{cls}: wipp_kernel<{A}, G=16> (implicit pivoting, two matrices per warp, L in tensor memory)"
        if n <= 32:
            return f"{cls}: wip_kernel<{A}, n={n}> (implicit pivoting, L in tensor memory, DMMA rank-8 updates)"
        return f"{cls}: wip2_kernel<{A}> (implicit pivoting, two rows per lane, L in tensor memory, DMMA updates)"
    if cls == "thread" and n > 8:
        return f"{cls}: tps_kernel<{A}, N={n}> (one thread per matrix, lower triangle in thread-private shared memory)"
    if cls == "thread":
        return f"{cls}: tpr_kernel<{A}, N={n}> (one thread per matrix, lower triangle in registers, tile staged through shared memory)"
    if cls == "cta_blocked":
        return f"{cls}: {'bsm_gmw_kernel' if algo == 'gmw81' else 'bsm_se_kernel'}<{A}> (panel in shared memory, DMMA trailing update from L2)"
    names = {"warp_smem": "wsm_gmw_kernel" if algo == "gmw81" else "wsm_se_kernel",
             "cta_packed": "csm_gmw_kernel" if algo == "gmw81" else "csm_se_kernel",
             "cta_smem": "gmw81_cta_kernel" if algo == "gmw81" else "se_cta_kernel",
             "cta_gmem": "gmw81_cta_kernel" if algo == "gmw81" else "se_cta_kernel", "warp32": "warp_*_kernel"}
    return f"{cls}: {names.get(cls, cls)}<{A}, {mode}>"


def time_steps(torch, fn, steps, warmup):
    """Device time of `steps` calls of fn() on torch's current stream, in ms per call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def other_configs(torch, np, m, dev, peak):
    """Short timed runs of BASELINE.json configs[0..3] and of the reference's bench family
    (rank 0, N = 1).  Inputs are resident; every batch is larger than the L2 or flushed by the
    alternation of input and output streams (stated per entry)."""
    import oracle
    from modcholesky_b200 import utils
    out = {}
    dfma, dmma = m.fp64_peak_tflops()
    out["fp64_peaks_tflops"] = {"dfma": dfma, "dmma_m8n8k4": dmma,
                                "how": "modchol_fp64_peak_tflops: dependent FMA / mma.sync chains, CUDA events"}

    def device_case(algo, n, batch, a, tag):
        ent = {"workload": tag, "algo": algo, "n": n, "batch": batch,
               "bytes_touched_per_step": batch * (16 * n * n + 16 * n)}
        for mode in ("fast", "strict"):
            ms = time_steps(torch, lambda: m.factor_batched(algo, a, mode=mode), 5, 3)
            rate = batch / (ms * 1e-3)
            gbs = rate * (16 * n * n + 16 * n) / 1e9
            ent[mode] = {"matrices_per_sec": rate, "ms_per_step": ms, "kernel": kernel_name(m, algo, n, mode),
                         "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                      "frac": gbs / peak}}
            if n >= 128:
                tf = rate * n ** 3 / 3.0 / 1e12
                ent[mode]["fp64"] = {"achieved_tflops": tf, "frac_of_dfma_peak": tf / dfma,
                                     "frac_of_dmma_peak": tf / dmma, "flops_per_matrix": n ** 3 / 3.0}
        return ent

    # configs[1..3]: family F1 (symmetric uniform), as the headline
    for key, algo, n, batch in (("configs[1]", "se99", 16, 1 << 20), ("configs[2]", "se90", 64, 200_000),
                                ("configs[3]", "gmw81", 256, 4096)):
        a = gen_batch(torch, batch, n, SEED + 17 * n, dev)
        ent = device_case(algo, n, batch, a, f"{algo} n={n}, {batch} symmetric-uniform indefinite matrices, resident")
        d = m.factor_batched(algo, a, mode="fast")
        resid, flags = m.verify_batched(a, d)
        amax = a.abs().amax(dim=(1, 2))
        bar = torch.clamp(8.0 * 2.220446049250313e-16 * (amax + d.e.amax(dim=1)) / torch.clamp(amax, min=1.0), min=1e-12)
        ent["verified_on_device"] = bool(int(flags.max()) == 0 and bool((resid <= bar).all()))
        if key == "configs[1]":
            vms = time_steps(torch, lambda: m.verify_batched(a, d), 3, 1)
            ent["verify_ms"] = vms
            # the other kernels either side of the path, on the same batch, against the HBM roofline
            gms = time_steps(torch, lambda: m.gershgorin_circles_batched(a), 3, 1)
            gb = batch * (8 * n * n + 16 * n) / (gms * 1e-3) / 1e9
            vb = batch * (16 * n * n + 16 * n + 12) / (vms * 1e-3) / 1e9
            rhs = torch.ones((batch, n), dtype=torch.float64, device=dev)
            sms_ = time_steps(torch, lambda: m.solve_batched(d, rhs), 3, 1)
            sb = batch * (8 * n * n + 24 * n) / (sms_ * 1e-3) / 1e9
            ent["aux_kernels"] = {
                "gershgorin_circles": {"ms": gms, "GBps": gb, "frac_of_hbm_peak": gb / peak,
                                       "algorithmic_bytes_per_matrix": 8 * n * n + 16 * n},
                "verify": {"ms": vms, "GBps": vb, "frac_of_hbm_peak": vb / peak,
                           "algorithmic_bytes_per_matrix": 16 * n * n + 16 * n + 12},
                "solve": {"ms": sms_, "GBps": sb, "frac_of_hbm_peak": sb / peak,
                          "algorithmic_bytes_per_matrix": 8 * n * n + 24 * n}}
            del rhs
        out[key] = ent
        del a, d, resid, flags
        torch.cuda.empty_cache()

    # the reference's fixed-size bench shapes 3x3 / 4x4 (benches/bench.rs:23-44 and the se90 / se99
    # twins), batched: 2^22 symmetric-uniform matrices through the thread-per-matrix kernels; n = 8
    # (2^20 matrices) is the upper end of the register-resident class
    small = {}
    for n in (3, 4, 8):
        batch = 1 << 22 if n <= 4 else 1 << 20
        a = gen_batch(torch, batch, n, SEED + 17 * n, dev)
        ent = {}
        for algo in ("gmw81", "se90", "se99"):
            c = device_case(algo, n, batch, a, f"{algo} n={n}")
            ent[algo] = {"fast": {k: c["fast"][k] for k in ("matrices_per_sec", "ms_per_step", "kernel")},
                         "fast_roofline_frac": c["fast"]["roofline"]["frac"],
                         "strict_matrices_per_sec": c["strict"]["matrices_per_sec"],
                         "strict_roofline_frac": c["strict"]["roofline"]["frac"]}
        ah = a[:4096].cpu().numpy()
        ds = m.factor_batched("se99", a[:4096], mode="fast")
        ref = oracle.factor_batched("se99", ah, threads=0)
        ent["fast_bit_exact_vs_oracle"] = bool(np.array_equal(ds.l.cpu().numpy(), ref.l)
                                               and np.array_equal(ds.e.cpu().numpy(), ref.e)
                                               and np.array_equal(ds.p.cpu().numpy().astype(np.uint64), ref.p))
        ent["batch"] = batch
        small[f"n={n}"] = ent
        del a
        torch.cuda.empty_cache()
    out["small_n_thread_per_matrix"] = small

    # the reference's bench family on its own shapes (benches/bench.rs:46-83): Q D Q^T with the
    # bit-exact seeded factors, assembled on the device; the CPU column is the oracle on the
    # same bytes (all host threads)
    fam = {}
    for n, batch, cpu_n in ((12, 1 << 18, 1 << 14), (128, 8192, 256), (256, 4096, 64)):
        v, dg = utils.bench_input_batch(n, batch)
        vd, dd = torch.from_numpy(v).to(dev), torch.from_numpy(dg).to(dev)
        asm_ms = time_steps(torch, lambda: m.assemble_qdqt_batched(vd, dd), 2, 1)
        a = m.assemble_qdqt_batched(vd, dd)
        ent = {"batch": batch, "assemble_ms": asm_ms, "algos": {}}
        ah = a[:cpu_n].cpu().numpy()
        for algo in ("gmw81", "se90", "se99"):
            c = device_case(algo, n, batch, a, f"bench.rs {algo}_{n}x{n}_nd family")
            t0 = time.perf_counter()
            ref = oracle.factor_batched(algo, ah, threads=0)
            dt = time.perf_counter() - t0
            ds = m.factor_batched(algo, a[:cpu_n], mode="strict")
            same = bool(np.array_equal(ds.l.cpu().numpy(), ref.l) and np.array_equal(ds.e.cpu().numpy(), ref.e)
                        and np.array_equal(ds.p.cpu().numpy().astype(np.uint64), ref.p))
            ent["algos"][algo] = {"gpu_fast_matrices_per_sec": c["fast"]["matrices_per_sec"],
                                  "gpu_strict_matrices_per_sec": c["strict"]["matrices_per_sec"],
                                  "gpu_fast_roofline_frac": c["fast"]["roofline"]["frac"],
                                  "kernel_fast": c["fast"]["kernel"],
                                  "cpu_oracle_matrices_per_sec": cpu_n / dt, "cpu_threads": ref.threads,
                                  "cpu_sample": cpu_n, "strict_bit_exact_vs_oracle": same}
        fam[f"n={n}"] = ent
        del a, vd, dd
        torch.cuda.empty_cache()
    out["reference_bench_family"] = fam

    # configs[0]: the reference's single-matrix case -- one 10x10 / 12x12 Q D Q^T matrix through
    # the three algorithms on the CPU (oracle, one thread, us per call) and the latency of the same
    # call through the trait path on the GPU (batch = 1: HOST buffers, and DEVICE buffers)
    c0 = {}
    for n in (10, 12):
        a1 = utils.bench_input_dense(n)
        ent = {}
        for algo in ("gmw81", "se90", "se99"):
            reps = 2000
            stack = np.ascontiguousarray(np.broadcast_to(a1, (reps, n, n)))
            oracle.factor_batched(algo, stack[:16], threads=1)
            t0 = time.perf_counter()
            oracle.factor_batched(algo, stack, threads=1)
            cpu_us = (time.perf_counter() - t0) / reps * 1e6
            arr = m.Array2(a1, mode="strict")
            fn = getattr(arr, "mod_cholesky_" + algo)
            for _ in range(20):
                fn()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            host_us = (time.perf_counter() - t0) / 200 * 1e6
            ad = torch.from_numpy(a1[None]).to(dev)
            for _ in range(20):
                m.factor_batched(algo, ad)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                m.factor_batched(algo, ad)
            torch.cuda.synchronize()
            dev_us = (time.perf_counter() - t0) / 200 * 1e6
            kern_us = time_steps(torch, lambda: m.factor_batched(algo, ad), 50, 5) * 1e3
            ent[algo] = {"cpu_oracle_us_per_call": cpu_us, "gpu_trait_call_us_host_buffers": host_us,
                         "gpu_call_us_device_buffers": dev_us, "gpu_device_time_us": kern_us}
        c0[f"{n}x{n}"] = ent
    c0["note"] = ("benches/bench.rs has 12x12 (bench.rs:46-57), BASELINE.json says 10x10: both; CPU = C oracle, "
                  "one thread, 2000 calls; GPU = Array2(...).mod_cholesky_*() with numpy input (H2D + kernel + "
                  "D2H per call) and the same call on resident torch tensors (wall clock incl. Python)")
    out["configs[0]"] = c0
    return out


def cpu_baseline(a_host, threads=0):
    import oracle
    oracle.build_oracle()
    t0 = time.perf_counter()
    r = oracle.factor_batched(ALGO, a_host, threads=threads)
    dt = time.perf_counter() - t0
    return a_host.shape[0] / dt, r.threads, dt


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path.  The crate is
    Rust and cannot be compiled in this image (no rustc/cargo), so this times the oracle port
    (oracle/modchol_oracle.c, a literal C restatement) looped over the batch on all host cores
    -- the north-star's "ndarray path looped with rayon" stand-in."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    sample_src = "numpy default_rng"
    a = None
    try:   # the same bytes as the first CPU_SAMPLE matrices of the CUDA arm's batch on rank 0
        import torch
        if torch.cuda.is_available():
            a = gen_batch(torch, CPU_SAMPLE, N_MAT, SEED, torch.device("cuda", 0)).cpu().numpy()
            sample_src = "the first matrices of the CUDA arm's rank-0 batch (same generator, same seed)"
            torch.cuda.empty_cache()
    except Exception:
        a = None
    if a is None:
        rng = np.random.default_rng(SEED)
        a = rng.uniform(-1, 1, size=(CPU_SAMPLE, N_MAT, N_MAT))
        a = np.tril(a) + np.transpose(np.tril(a, -1), (0, 2, 1))
    for _ in range(max(args.warmup, 1)):
        cpu_baseline(a[: CPU_SAMPLE // 8])
    t0 = time.perf_counter()
    threads = 1
    for _ in range(args.steps):
        _, threads, _ = cpu_baseline(a)
    dt = time.perf_counter() - t0
    value = args.steps * CPU_SAMPLE / dt
    line = {
        "impl": "reference", "metric": "matrices_per_sec", "value": value, "unit": "matrices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"se99 n={N_MAT} symmetric-uniform, {CPU_SAMPLE} matrices per step "
                               f"(bounded sample of the 2^20-per-GPU workload)", "algo": ALGO, "n": N_MAT},
        "cpu_baseline": {"value": value, "unit": "matrices/s", "cores": threads, "kind": "port",
                         "sample": f"{CPU_SAMPLE} matrices x {args.steps} steps ({sample_src}), C restatement of "
                                   f"the reference (Rust toolchain absent), pthreads over the batch"},
        "e2e": {"value": value, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import modcholesky_b200 as m
    from modcholesky_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = args.batch
    a = gen_batch(torch, batch, N_MAT, SEED + rank, dev)
    stream = torch.cuda.current_stream()

    def step(mode):
        return m.factor_batched(ALGO, a, mode=mode)

    # ---- device-resident throughput ------------------------------------------------------
    def timed(mode, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            d = step(mode)
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = m.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            d = step(mode)
        e1.record(stream)
        barrier()
        launches = m.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches, clocks, d

    ms, launches, clocks, d = timed(args.mode, args.steps, args.warmup, sample_clocks=True)
    ms_per_step = ms / args.steps
    value = world * batch / (ms_per_step * 1e-3)
    other = "strict" if args.mode == "fast" else "fast"
    ms_o, _, _, _ = timed(other, max(2, args.steps // 2), 3)
    value_other = world * batch / (ms_o / max(2, args.steps // 2) * 1e-3)

    # parity spot check against the oracle on rank 0 (checker only; not timed)
    parity = None
    if rank == 0:
        import oracle
        sub = a[:512].cpu().numpy()
        ref = oracle.factor_batched(ALGO, sub, threads=0)
        ds = m.factor_batched(ALGO, a[:512], mode="strict")
        parity = bool(np.array_equal(ds.l.cpu().numpy(), ref.l) and np.array_equal(ds.e.cpu().numpy(), ref.e)
                      and np.array_equal(ds.p.cpu().numpy().astype(np.uint64), ref.p))
        resid, flags = m.verify_batched(a, d)
        parity = parity and int(flags.max()) == 0 and float(resid.max()) <= 1e-12
    del d

    # ---- end to end through the public API with HOST (pinned) buffers -----------------------
    eb = min(E2E_BATCH_PER_GPU, batch)
    ah = torch.empty((eb, N_MAT, N_MAT), dtype=torch.float64, pin_memory=True)
    ah.copy_(a[:eb])
    lh = torch.empty_like(ah, pin_memory=True)
    eh = torch.empty((eb, N_MAT), dtype=torch.float64, pin_memory=True)
    ph = torch.empty((eb, N_MAT), dtype=torch.int64, pin_memory=True)
    L = _lib.lib()
    mode_id = _lib.mode_id(args.mode)

    def e2e_step():
        rc = L.modchol_factor_batched_f64(_lib.SE99, mode_id, eb, N_MAT, ah.data_ptr(), lh.data_ptr(),
                                          eh.data_ptr(), ph.data_ptr(), None, _lib.MEM_HOST, None)
        _lib.check(rc)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                       # returns when the results are in host memory
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * eb * e2e_steps / float(dt.item())
    # the trait path a reference user takes: pageable numpy input through Array3(...), outputs
    # allocated by the call (the library bounces through cached pinned buffers); rank 0, N = 1
    e2e_pageable = None
    if rank == 0 and world == 1:
        an = ah.numpy().copy()
        arr = m.Array3(an, mode=args.mode)
        arr.mod_cholesky_se99_batched()
        t0 = time.perf_counter()
        for _ in range(2):
            arr.mod_cholesky_se99_batched()
        e2e_pageable = 2 * eb / (time.perf_counter() - t0)
        del an, arr
    h2d = eb * N_MAT * N_MAT * 8
    d2h = eb * (N_MAT * N_MAT * 8 + N_MAT * 8 + N_MAT * 8)

    # ---- roofline of the dominant (only) kernel -------------------------------------------
    peak, peak_src = measured_peak()
    achieved = batch * ALG_BYTES / (ms_per_step * 1e-3) / 1e9      # per GPU: one launch per step
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(),
                "kernel": kernel_name(m, ALGO, N_MAT, args.mode) + " (one launch per step)",
                "algorithmic_bytes_per_matrix": ALG_BYTES, "peak_source": peak_src,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu capture (profiles/r02_traffic.json), scaled to this batch",
                "note": "latency / issue bound (one warp per matrix, 32 dependent pivot steps), not HBM bound: see profiles/r02_notes.md"}

    line = {
        "metric": "matrices_per_sec", "value": value, "unit": "matrices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"se99 n={N_MAT}, {batch} symmetric-uniform indefinite matrices per GPU "
                               f"(BASELINE.json configs[4] per-GPU slice)",
                   "algo": ALGO, "n": N_MAT, "batch_per_gpu": batch, "mode": args.mode,
                   "family": "F1 symmetric-uniform (phase 2 from k=0)",
                   "l2": "inputs (8.6 GB) + outputs (8.9 GB) per step exceed the 126 MB L2; no flush needed",
                   "sharding": "independent contiguous batch slices, one process per GPU, no collective"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "batch_per_gpu": eb, "steps": e2e_steps,
                "trait_path_pageable_numpy": e2e_pageable,
                "host_ceiling": "profiles/r02_pcie_ceiling_{1,8}gpu.txt: pinned H2D+D2H at once tops out at 95.5 GB/s "
                                "for one rank and 127.7 GB/s for eight (all GPUs behind NUMA node 0): 5.65 / 7.56 M "
                                "matrices/s at 16 896 B per matrix",
                "path": "modchol_factor_batched_f64(mem_kind=HOST): pinned host -> 3-slot overlapped "
                        "H2D / kernel / D2H pipeline -> pinned host"},
        "roofline": roofline,
        f"value_{other}_mode": value_other,
        "parity_spot_check": parity,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sub = a[:CPU_SAMPLE].cpu().numpy()
        v, cores, secs = cpu_baseline(sub)
        line["cpu_baseline"] = {"value": v, "unit": "matrices/s", "cores": cores, "kind": "port",
                                "sample": f"{sub.shape[0]} of the {batch} matrices, {secs:.1f} s wall on "
                                          f"{cores} threads (C restatement of the reference; Rust absent)"}
    # the thread-per-device host entry point (modchol_factor_batched_multi_gpu_f64) on every GPU of the
    # job: rank 0, host buffers, untimed, checked against the single-device result
    if world > 1:
        if rank == 0:
            sub = a[:4096].cpu().numpy()
            d1 = m.factor_batched(ALGO, sub, mode="strict")
            dn = m.factor_batched(ALGO, sub, mode="strict", ngpus=world)
            line["multi_gpu_entry_check"] = bool(np.array_equal(d1.l, dn.l) and np.array_equal(d1.e, dn.e)
                                                 and np.array_equal(d1.p, dn.p))
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_configs:
        del a
        torch.cuda.empty_cache()
        line["configs"] = other_configs(torch, np, m, dev, peak)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
