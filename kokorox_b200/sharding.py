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


class LeastLoadedRouter:
    """Online counterpart of ``shard_requests`` for a server process that owns one session per GPU (SURVEY.md 8e:
    "route requests/chunks to GPUs by least outstanding audio-frames").  ``sessions`` are objects with
    ``infer_one(tokens, style, speed)`` (``B200Koko``); the estimate of a request's work is its token count, which
    is what the frame count is proportional to before the durations are known.  Thread-safe: the server's request
    threads call ``infer_one`` concurrently; with ``set_option("coalesce", K)`` on each session the callers routed
    to one GPU leave as ragged batches.  Chunks of one text must be re-concatenated by the caller in submission
    order (koko.rs:1179) -- the router does not reorder, it only picks the device."""

    def __init__(self, sessions):
        import threading
        if not sessions:
            raise ValueError("LeastLoadedRouter needs at least one session")
        self._sessions = list(sessions)
        self._load = [0] * len(self._sessions)
        self._served = [0] * len(self._sessions)
        self._lock = threading.Lock()

    def pick(self, cost: int) -> int:
        with self._lock:
            r = min(range(len(self._load)), key=lambda i: (self._load[i], i))
            self._load[r] += cost
            return r

    def done(self, r: int, cost: int) -> None:
        with self._lock:
            self._load[r] -= cost
            self._served[r] += 1

    def infer_one(self, tokens, style, speed: float = 1.0):
        cost = max(len(tokens), 1)
        r = self.pick(cost)
        try:
            return self._sessions[r].infer_one(tokens, style, speed)
        finally:
            self.done(r, cost)

    @property
    def outstanding(self):
        with self._lock:
            return list(self._load)

    @property
    def served(self):
        with self._lock:
            return list(self._served)
