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


def _hyper_engine():
    """A device-less engine over a small synthetic batch: the hyper-parameter sweep is host code, and elements whose chains never
    ran contribute their INIT_* rates (bpp.cpp:215-220), so the N > 1 exchange can be exercised on CPU."""
    import phyloacc_b200 as pab
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=10, seed=5, n_out=2, n_target=2)
    batch = synth.simulate(tree, C=9, length=60, seed=2)
    prm = pab.default_params()
    prm.seed = 17
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                            conserve=tree.conserve, device=-1)
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    return eng, batch.C


def test_hyper_sums_of_shards_add_up(native_lib):
    eng, C = _hyper_engine()
    ids = np.arange(C, dtype=np.int32)
    total = eng.hyper_sums(ids)
    parts = sum(eng.hyper_sums(ids[list(r)]) for r in shard.shard_ranges(np.ones(C), 3))
    assert total[10] == C and np.allclose(total, parts, rtol=1e-14, atol=0)
    before = eng.hyper()
    want = eng.sample_hyper(4, ids)
    eng.set_hyper(**before)
    assert eng.sample_hyper_from_sums(4, total) == want
    assert shard.sample_hyper_sharded(eng, 5, ids, world=1) == eng.hyper()
    eng.close()


def _hyper_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng, C = _hyper_engine()
    mine = np.array(list(shard.shard_ranges(np.ones(C), world)[rank]), dtype=np.int32)
    out = [shard.sample_hyper_sharded(eng, it, mine, world) for it in range(3)]  # three outer CHAIN iterations
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def test_sample_hyper_sharded_world2_gloo(native_lib):
    """Two ranks, each holding half of the elements: one all-reduce of 11 sums per sweep, the same hyper-parameters on both ranks,
    equal to the single-process sweep over all elements."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hyper_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    eng, C = _hyper_engine()
    ids = np.arange(C, dtype=np.int32)
    want = [eng.sample_hyper(it, ids) for it in range(3)]
    eng.close()
    assert got[0] == got[1]
    for a, b in zip(got[0], want):
        assert a.keys() == b.keys() and all(abs(a[k] - b[k]) <= 1e-12 * abs(b[k]) for k in a)
    assert want[2] != want[0] or want[1] != want[0]
