"""Host-side logic of the multi-GPU driver on CPU: symbol sharding (no collective on the data path)
and the bench's barrier + max-over-ranks timing reduction, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

from polars_quant_b200 import shard


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 500, 5000, 10_000, 50_000])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_ranges_partition_the_panel_in_whole_blocks(n, world):
    rs = shard.all_ranges(n, world)
    assert rs[0][0] == 0 and rs[-1][1] == n
    for (a, b), (c, d) in zip(rs, rs[1:]):
        assert b == c and a <= b
    for lo, hi in rs:
        assert lo % shard.BLOCK == 0 or lo == n
    sizes = [-(-(hi - lo) // shard.BLOCK) for lo, hi in rs]
    assert max(sizes) - min(sizes) <= 1                      # balanced to one block


def test_shard_columns_keeps_panel_order():
    names = [f"S{i:05d}" for i in range(100)]
    got = sum((shard.shard_columns(names, 4, r) for r in range(4)), [])
    assert got == names


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_symbols, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.symbol_range(n_symbols, world, rank)
    # every rank "processes" its own range: mark ownership, then check global coverage with a
    # reduction that is NOT on the data path of the product (test-only)
    owned = torch.zeros(n_symbols, dtype=torch.int32)
    owned[lo:hi] = 1
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    # the bench's timing reduction: barrier, then MAX over ranks of the per-rank device time
    dist.barrier()
    t = torch.tensor([10.0 + rank, 9.0 + 2 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    q.put((rank, lo, hi, bool((owned == 1).all()), t.tolist(), float(units[0])))
    dist.destroy_process_group()


def test_two_ranks_cover_the_panel_once_and_reduce_timings():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 5000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, cov0, t0, u0), (r1, lo1, hi1, cov1, t1, u1) = res
    assert (lo0, hi1) == (0, n) and hi0 == lo1 and hi0 % shard.BLOCK == 0
    assert cov0 and cov1                                     # every symbol owned exactly once
    assert t0 == t1 == [11.0, 11.0]                          # max over ranks
    assert u0 == u1 == float(n)
