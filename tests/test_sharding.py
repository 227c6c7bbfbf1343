"""CPU tests for the N>1 path (world_size-2 gloo): request sharding and the bench reductions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kokorox_b200.sharding import reduce_step, shard_requests


def test_shard_requests_partitions_and_balances():
    rng = np.random.default_rng(4)
    costs = rng.integers(10, 511, 4096).tolist()           # BASELINE configs[4]: N ~ U{10..510}
    for world in (1, 2, 4, 8):
        sh = shard_requests(costs, world)
        flat = sorted(i for s in sh for i in s)
        assert flat == list(range(len(costs)))              # every request exactly once
        assert all(s == sorted(s) for s in sh)              # in-rank order preserved
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= 510               # within one request of perfect balance
    assert shard_requests([], 2) == [[], []]
    assert shard_requests([5], 4) == [[0], [], [], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = list(range(10, 110))
    mine = shard_requests(costs, world)[rank]
    step_s = 0.5 + 0.25 * rank                               # rank 1 is slower
    audio_s = float(sum(costs[i] for i in mine))
    dist.barrier()
    t, (a, n) = reduce_step(step_s, [audio_s, float(len(mine))])
    q.put((rank, mine, t, a, n))
    dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    all_idx = sorted(res[0][1] + res[1][1])
    assert all_idx == list(range(100))
    for _, _, t, a, n in res:
        assert t == 0.75                                      # max over ranks
        assert a == float(sum(range(10, 110))) and n == 100.0  # summed over ranks


def test_reduce_step_single_process():
    assert reduce_step(1.5, [2.0, 3.0]) == (1.5, [2.0, 3.0])


def test_least_loaded_router_balances_concurrent_requests():
    # online routing by least outstanding work, one session per GPU, request threads in one process
    import threading
    import time
    from kokorox_b200.sharding import LeastLoadedRouter

    class FakeSession:
        def __init__(self):
            self.calls, self.active, self.peak = [], 0, 0
            self.lock = threading.Lock()

        def infer_one(self, tokens, style, speed):
            with self.lock:
                self.active += 1
                self.peak = max(self.peak, self.active)
            time.sleep(0.0005 * len(tokens))          # "GPU time" proportional to the token count
            with self.lock:
                self.active -= 1
                self.calls.append(len(tokens))
            return len(tokens)

    sessions = [FakeSession() for _ in range(4)]
    router = LeastLoadedRouter(sessions)
    rng = np.random.default_rng(0)
    lens = rng.integers(10, 200, size=64)
    out = [None] * len(lens)

    def client(i):
        out[i] = router.infer_one([0] * int(lens[i]), None, 1.0)
    threads = [threading.Thread(target=client, args=(i,)) for i in range(len(lens))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert out == [int(n) for n in lens]                      # every caller got its own result
    assert router.outstanding == [0, 0, 0, 0] and sum(router.served) == len(lens)
    work = [sum(s.calls) for s in sessions]
    assert min(work) > 0.6 * max(work), work                 # token work spread over the four devices
    with pytest.raises(ValueError):
        LeastLoadedRouter([])
    # a failing session releases its load
    class Boom:
        def infer_one(self, *a):
            raise RuntimeError("x")
    r2 = LeastLoadedRouter([Boom()])
    with pytest.raises(RuntimeError):
        r2.infer_one([0, 1, 0], None, 1.0)
    assert r2.outstanding == [0]
