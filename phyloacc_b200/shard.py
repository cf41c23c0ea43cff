"""Multi-GPU plumbing for the element loop: elements are independent, so the ids of a run are split into contiguous,
cost-balanced ranges (one per rank / GPU) with NO data-path collective; the only communication is one gather of the
per-element summaries at the end of the run.  torch.distributed provides the plumbing (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_ranges(costs: Sequence[float], world: int) -> List[range]:
    """Contiguous ranges over len(costs) items whose cost sums are as even as a prefix split allows.
    cost ~ GG * (S - 1) per element."""
    n = len(costs)
    c = np.asarray(costs, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [range(bounds[r], bounds[r + 1]) for r in range(world)]


def gather_summaries(local: np.ndarray, counts: Sequence[int], rank: int, world: int, device=None) -> np.ndarray:
    """All-gather per-element summary rows ([n_local, k] float64) of every rank into [sum(counts), k], in rank order.
    `counts[r]` is the number of rows rank r contributes (known to every rank from shard_ranges)."""
    import torch
    import torch.distributed as dist

    k = local.shape[1] if local.ndim == 2 else 1
    local = np.ascontiguousarray(local, dtype=np.float64).reshape(-1, k)
    if world == 1:
        return local
    m = max(counts)
    buf = torch.zeros((m, k), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([out[r][: counts[r]].cpu().numpy() for r in range(world)], axis=0)


def sample_hyper_sharded(engine, it: int, local_ids, world: int, device=None) -> dict:
    """BPP::sample_hyperparam (bpp.cpp:324-477) with the elements sharded over ranks: the sweep depends on the elements only
    through 11 sums (`Engine.hyper_sums`), so each rank adds up its own ids, ONE all-reduce adds the ranks, and every rank runs
    the same sweep on the totals (same Philox draws -> same accepted values everywhere).  SURVEY.md §8e."""
    import torch
    import torch.distributed as dist

    sums = engine.hyper_sums(local_ids)
    if world > 1:
        t = torch.from_numpy(sums.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sums = t.cpu().numpy()
    return engine.sample_hyper_from_sums(it, sums)
