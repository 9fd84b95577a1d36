"""Multi-GPU sharding of a batch.

The matrices are independent, so the hot path has NO collective: rank r of `world` factorises
the contiguous slice [lo, hi) of the batch on its own GPU.  The split is the one the C-ABI's
modchol_factor_batched_multi_gpu_f64 uses (per = ceil(batch / world)).  The only exchange is
the OPTIONAL gather of the results (NCCL all_gather over NVLink on GPUs, gloo on CPU tests),
which is outside every timed region.
"""
from __future__ import annotations

from typing import Tuple

from .api import Decomposition


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of rank `rank`; identical to the slices of modchol_api.cu."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def gather_decomposition(local: Decomposition, batch: int, group=None) -> Decomposition:
    """All-gather the per-rank slices of a sharded Decomposition (torch tensors, CPU with gloo or
    CUDA with NCCL) back into full-batch tensors on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = (batch + world - 1) // world

    def gather(t):
        pad_shape = (per,) + tuple(t.shape[1:])
        padded = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        padded[: t.shape[0]] = t
        out = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(out, padded, group=group)
        return torch.cat(out, dim=0)[:batch]

    k = gather(local.phase2_start) if local.phase2_start is not None else None
    return Decomposition(gather(local.l), gather(local.e), gather(local.p), k)
