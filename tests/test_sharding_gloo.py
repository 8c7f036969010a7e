"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, global seeding, rollout gather,
global advantage statistics, episode-stat reduction, table broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from steppingstone_b200 import seeding, sharding


def test_shard_bounds_partition():
    for n, w in ((65536, 8), (8192, 1), (10, 3), (7, 8)):
        covered = []
        for r in range(w):
            first, k = sharding.shard_bounds(r, w, n)
            covered += list(range(first, first + k))
        assert covered == list(range(n))


def test_shard_seeds_keep_global_numbering():
    full = seeding.env_streams(8, 0, 6)
    for r in range(2):
        first, k = sharding.shard_bounds(r, 2, 6)
        assert np.array_equal(seeding.env_streams(8, first, k), full[first:first + k])
        assert np.array_equal(seeding.aux_streams(1, 8, first, k), seeding.aux_streams(1, 8, 0, 6)[first:first + k])


class _FakeRollouts:
    pass


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, N = 5, 8
        g = torch.Generator().manual_seed(0)
        full = {f: torch.randn(T + 1, N, 3, generator=g) for f in sharding.ROLLOUT_FIELDS}
        first, k = sharding.shard_bounds(rank, world, N)
        ro = _FakeRollouts()
        for f in sharding.ROLLOUT_FIELDS:
            setattr(ro, f, full[f][:, first:first + k].clone())
        gathered = sharding.gather_rollout(ro)
        ok = all(torch.equal(gathered[f], full[f]) for f in sharding.ROLLOUT_FIELDS)
        adv = full["returns"][:-1] - full["value_preds"][:-1]
        mean, std = sharding.global_advantage_stats(adv[:, first:first + k])
        ok = ok and abs(float(mean) - float(adv.mean())) < 1e-6 and abs(float(std) - float(adv.std())) < 1e-6
        m, c = sharding.reduce_episode_stats(torch.tensor(10.0 * (rank + 1)), torch.tensor(rank + 1))
        ok = ok and c == 3 and abs(m - 10.0) < 1e-9
        table = torch.full((4,), float(rank))
        sharding.broadcast_table(table, src=0)
        ok = ok and bool((table == 0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_collectives():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
