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
