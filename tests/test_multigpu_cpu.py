"""N > 1 host logic on CPU: world_size-2 `gloo` process group (no GPU): the equation /
draw partition every rank derives must tile the work exactly, and the rendezvous plumbing
bench.py uses (broadcast of the 128-byte communicator id from rank 0) must deliver the same
bytes to every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesianvars_b200.bvar import shard_draws, shard_equations


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, R, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    j0, n = shard_equations(M, rank, world)
    r0, nr = shard_draws(R, rank, world)
    mine = torch.zeros(M + R, dtype=torch.int32)
    mine[j0:j0 + n] += 1
    mine[M + r0:M + r0 + nr] += 1
    dist.all_reduce(mine)            # every equation / draw owned exactly once
    # the in-place all-gather layout: rank r's block starts at r * ceil(M / world) columns
    block = -(-M // world)
    cols = torch.full((world * block,), -1, dtype=torch.int64)
    cols[rank * block: rank * block + n] = torch.arange(j0, j0 + n)
    gathered = [torch.empty(block, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, cols[rank * block:(rank + 1) * block].clone())
    flat = torch.cat(gathered)[:M]
    if rank == 0:
        out.put((bool((mine == 1).all()), bool(torch.equal(flat, torch.arange(M))), uid.tolist() == list(range(128))))
    else:
        assert uid.tolist() == list(range(128))
    dist.destroy_process_group()


def test_world2_partition_and_id_broadcast():
    world = 2
    for M, R in [(100, 10000), (13, 7), (1, 1)]:
        ctx = mp.get_context("spawn")
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, M, R, q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get() == (True, True, True)


def test_shard_helpers_edge_cases():
    for M in (1, 7, 100, 200):
        for world in (1, 2, 4, 8):
            owned = np.zeros(M, int)
            for r in range(world):
                j0, n = shard_equations(M, r, world)
                owned[j0:j0 + n] += 1
            assert (owned == 1).all()
    assert shard_equations(3, 7, 8) == (3, 0)
