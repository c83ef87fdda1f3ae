"""N>1 host logic on CPU: round-robin page sharding + the equal-size all-gather used to bring results back (gloo,
world size 2, spawned processes).  The NCCL/NVLink variant of the same code runs in bench.py under torchrun."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mit_b200.pipeline import shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_pages, rank, world)
    # stand-in "page results": each page's buffer is filled with its global page index
    local = torch.stack([torch.full((4, 6), float(i)) for i in mine])
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local)
    gathered = torch.stack(out)                      # [world, pages_per_rank, ...]
    # rank 0 de-interleaves back to page order: page i lives at [i % world, i // world]
    order = torch.stack([gathered[i % world, i // world] for i in range(n_pages)])
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)         # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        q.put((order[:, 0, 0].tolist(), t.item(), mine))
    dist.destroy_process_group()


def test_round_robin_shard_and_gather_world2():
    assert shard_indices(8, 0, 2) == [0, 2, 4, 6] and shard_indices(8, 1, 2) == [1, 3, 5, 7]
    assert sorted(sum((shard_indices(256, r, 8) for r in range(8)), [])) == list(range(256))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    order, tmax, mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert order == [float(i) for i in range(8)] and tmax == 2.0 and mine == [0, 2, 4, 6]
