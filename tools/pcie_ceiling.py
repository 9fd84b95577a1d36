"""Host <-> device copy ceiling of this box, measured the way the e2e path uses it: every rank
(one per GPU, torchrun) copies pinned host buffers to its GPU and back, alone and with all ranks
at once.  Prints per-rank and aggregate GB/s, the GPU topology and the ranks' CPU affinity --
the evidence behind DESIGN.md section 7 (why the end-to-end number does not scale with GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_ceiling.py [--mib 1024]
"""
import argparse
import os
import subprocess
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

n = args.mib << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.ones(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind):
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = args.reps * n * (2 if kind == "both" else 1)
    return nbytes / dt / 1e9, world * nbytes / float(t.item()) / 1e9


if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
    except Exception as ex:
        print("nvidia-smi topo failed:", ex)
    print("host cores:", os.cpu_count(), " ranks:", world)
aff = sorted(os.sched_getaffinity(0))
for kind in ("h2d", "d2h", "both"):
    run(kind)
    mine, agg = run(kind)
    line = f"rank {rank}: {kind:4s} {mine:7.1f} GB/s (this rank, its own clock)   cpus {aff[0]}-{aff[-1]} ({len(aff)})"
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, line)
    else:
        out = [line]
    if rank == 0:
        print("\n".join(out))
        print(f"==> {kind}: aggregate over {world} rank(s), max-over-ranks clock: {agg:7.1f} GB/s")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
