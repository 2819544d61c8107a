"""N > 1 host logic on CPU: two gloo ranks (127.0.0.1) exercise the only collectives of the engine
(parameter broadcast, counter reductions) and the game sharding / per-rank seeding."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from matrix0_b200 import distributed as D


def test_shard_range_partitions_games():
    for total, world in [(32768, 8), (4096, 3), (7, 4), (0, 2)]:
        spans = [D.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert len({D.rank_seed(1234, r) for r in range(8)}) == 8


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)           # different weights on every rank before the broadcast
        params = {"b.weight": torch.randn(5, 3, generator=g), "a.bias": torch.randn(7, generator=g)}
        D.broadcast_parameters(params, src=0)
        g0 = torch.Generator().manual_seed(100)
        exp = {"b.weight": torch.randn(5, 3, generator=g0), "a.bias": torch.randn(7, generator=g0)}
        ok = all(torch.equal(params[k], exp[k]) for k in params)
        sims, ms = D.reduce_scalars([1000.0 * (rank + 1), 3.0]), D.reduce_scalars([10.0 + rank], op="max")
        start, stop = D.shard_range(10, rank, world)
        out.put((rank, ok, sims, ms, (start, stop)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_reductions():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sims, ms, span in res:
        assert ok                                   # every rank holds rank 0's parameters
        assert sims == [3000.0, 6.0] and ms == [11.0]
    assert res[0][4] == (0, 5) and res[1][4] == (5, 10)
