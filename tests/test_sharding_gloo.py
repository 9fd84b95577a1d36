"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch exactly like the C-ABI
multi-GPU entry point, each 'factorises' its slice (the CPU oracle stands in for the GPU here
-- test infrastructure only), and the optional gather reassembles the full result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modcholesky_b200.api import Decomposition
from modcholesky_b200.sharding import gather_decomposition, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 7, 8, 1000, (1 << 23)):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            assert max(hi - lo for lo, hi in spans) == (batch + world - 1) // world
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    rng = np.random.default_rng(99)          # every rank builds the same full batch
    a = rng.uniform(-1, 1, size=(batch, n, n))
    a = a + np.transpose(a, (0, 2, 1))
    lo, hi = shard_bounds(batch, world, rank)
    r = oracle.factor_batched("se99", a[lo:hi], threads=1)
    local = Decomposition(torch.from_numpy(r.l), torch.from_numpy(r.e),
                          torch.from_numpy(r.p.astype(np.int64)), torch.from_numpy(r.phase2_start))
    full = gather_decomposition(local, batch)
    ref = oracle.factor_batched("se99", a, threads=1)
    ok = (np.array_equal(full.l.numpy(), ref.l) and np.array_equal(full.e.numpy(), ref.e)
          and np.array_equal(full.p.numpy().astype(np.uint64), ref.p)
          and np.array_equal(full.phase2_start.numpy(), ref.phase2_start))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    world, batch, n = 2, 37, 9           # ragged: 19 + 18
    mp.spawn(_worker, args=(world, _free_port(), batch, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
