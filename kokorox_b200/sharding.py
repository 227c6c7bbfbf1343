"""Request sharding across the GPUs of one box.

Utterances (chunks) are independent -- the reference processes them one by one and concatenates
(/root/reference/kokorox/src/tts/koko.rs:947-1191) -- so the path shards by request with NO
data-path collective: one process + one kkx_ctx per GPU.  torch.distributed is used only for the
step barrier and for reducing the measured step time (max over ranks) and work (sum over ranks).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_requests(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy least-loaded assignment (longest first) of request indices to `world` ranks.
    `costs` = per-request work estimate (token count).  Deterministic; every rank computes the
    same table.  Within a rank, requests keep their original order (chunks of one text must be
    re-concatenated in order, koko.rs:1179)."""
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda r: (loads[r], r))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(x) for x in out]


def reduce_step(step_seconds: float, work: Sequence[float], device=None) -> Tuple[float, List[float]]:
    """(max over ranks of the step time, sum over ranks of each work counter).  Falls back to the
    local values when torch.distributed is not initialised (N=1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(step_seconds), [float(w) for w in work]
    t = torch.tensor([step_seconds], dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in w.tolist()]
