"""CPU tests of the N>1 path: contiguous cost-balanced element ranges per rank and the single gather of summaries,
with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phyloacc_b200 import shard


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    costs = rng.integers(50, 2000, size=1000).astype(float)
    for world in (1, 2, 3, 8):
        rs = shard.shard_ranges(costs, world)
        assert len(rs) == world
        assert [i for r in rs for i in r] == list(range(1000))
        sums = [costs[list(r)].sum() for r in rs]
        assert max(sums) - min(sums) <= 2 * costs.max()
    assert [len(r) for r in shard.shard_ranges([], 4)] == [0, 0, 0, 0]
    assert [list(r) for r in shard.shard_ranges([5.0], 2)] in ([[0], []], [[], [0]])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, costs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = shard.shard_ranges(costs, world)
    mine = np.array([[i, 10.0 * i] for i in rs[rank]], dtype=np.float64).reshape(-1, 2)
    allrows = shard.gather_summaries(mine, [len(r) for r in rs], rank, world)
    if rank == 0:
        q.put(allrows)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_summaries_world2_gloo():
    costs = [3.0, 1.0, 1.0, 1.0, 5.0, 1.0, 1.0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, costs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.array([[i, 10.0 * i] for i in range(len(costs))])
    assert np.array_equal(got, want)
