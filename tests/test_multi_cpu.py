"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: time-step sharding, level shards and
the one-level halo exchange that level-sharded z-derivatives need."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meteotools_b200.pipeline import halo_exchange_levels, level_shard, shard_steps


def test_shard_steps_partition():
    for n, w in ((96, 8), (10, 4), (3, 8), (1, 1)):
        got = sorted(s for r in range(w) for s in shard_steps(n, r, w))
        assert got == list(range(n))
        sizes = [len(shard_steps(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_level_shard_partition():
    for nz, w in ((80, 8), (60, 4), (7, 3), (5, 1)):
        cover = []
        for r in range(w):
            lo, hi, hlo, hhi = level_shard(nz, r, w)
            cover += list(range(lo, hi))
            assert hlo == (1 if r > 0 else 0) and hhi == (1 if r < w - 1 else 0)
        assert cover == list(range(nz))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, nz, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = [torch.arange(nz * 3 * 4, dtype=torch.float64).reshape(nz, 3, 4) * (k + 1) for k in range(2)]
        lo, hi, hlo, hhi = level_shard(nz, rank, world)
        local = []
        for f in full:
            buf = torch.full((hi - lo + hlo + hhi, 3, 4), -1.0, dtype=torch.float64)
            buf[hlo:hlo + hi - lo] = f[lo:hi]
            local.append(buf)
        halo_exchange_levels(local, hlo, hhi, rank, world, dist)
        ok = all(torch.equal(b, f[lo - hlo:hi + hhi]) for b, f in zip(local, full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, 9), (3, 7)])
def test_halo_exchange_gloo(world, nz):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, nz, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(res[r] for r in range(world)), res
