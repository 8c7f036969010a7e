"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, global seeding, rollout gather,
global advantage statistics, episode-stat reduction, table broadcast."""
import os
import socket

import numpy as np
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
        # rank-major (no env-axis transpose): [world, T+1, N_local, .] -- rank r's shard sits in block r
        rm = sharding.gather_rollout(ro, rank_major=True)
        ok = ok and all(tuple(rm[f].shape) == (world, T + 1, k, 3) for f in sharding.ROLLOUT_FIELDS)
        ok = ok and all(torch.equal(rm[f][r], full[f][:, r * k:(r + 1) * k]) for f in sharding.ROLLOUT_FIELDS for r in range(world))
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


def _fill_rollouts(ro, full, lo, hi):
    for f in sharding.ROLLOUT_FIELDS:
        getattr(ro, f).copy_(full[f][:, lo:hi])


def _dp_worker(rank, world, port, q):
    """Data-parallel PPO update (examples/train_ppo.py: global advantage statistics + averaged gradients) on two
    rollout shards == the reference-style update of the whole batch in one process."""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_ppo", os.path.join(root, "examples", "train_ppo.py"))
    tp = importlib.util.module_from_spec(spec); sys.modules["train_ppo"] = tp; spec.loader.exec_module(tp)
    from steppingstone_b200.storage import RolloutStorage
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, N = 6, 8
        g = torch.Generator().manual_seed(1)
        shapes = dict(observations=(T + 1, N, 60), states=(T + 1, N, 1), rewards=(T, N, 1), value_preds=(T + 1, N, 1),
                      returns=(T + 1, N, 1), action_log_probs=(T, N, 1), actions=(T, N, 21), masks=(T + 1, N, 1), bad_masks=(T + 1, N, 1))
        full = {f: torch.randn(*s, generator=g) for f, s in shapes.items()}
        full["actions"] = full["actions"] * 0.2
        full["action_log_probs"] = full["action_log_probs"] * 0.1 + 30.0

        def fresh(arch):
            torch.manual_seed(0)
            ac = tp.ActorCritic(60, 21, arch=arch).double()
            return ac, torch.optim.SGD(ac.parameters(), lr=1e-2)

        ok = True
        for arch in ("small", "reference"):
            first, k = sharding.shard_bounds(rank, world, N)
            ac, opt = fresh(arch)
            if rank == 1:                                  # rank 1 starts elsewhere: broadcast must bring it to rank 0's policy
                with torch.no_grad():
                    for p in ac.parameters():
                        p.add_(0.5)
            sharding.broadcast_parameters(ac)
            ro = RolloutStorage(T, k, (60,), 21, 1, device="cpu")
            for name in sharding.ROLLOUT_FIELDS:
                setattr(ro, name, getattr(ro, name).double())
            _fill_rollouts(ro, {f: v.double() for f, v in full.items()}, first, first + k)
            tp.ppo_update(ac, opt, ro, 1, 1, distributed=True)
            ac1, opt1 = fresh(arch)
            ro1 = RolloutStorage(T, N, (60,), 21, 1, device="cpu")
            for name in sharding.ROLLOUT_FIELDS:
                setattr(ro1, name, getattr(ro1, name).double())
            _fill_rollouts(ro1, {f: v.double() for f, v in full.items()}, 0, N)
            tp.ppo_update(ac1, opt1, ro1, 1, 1, distributed=False)
            err = max(float((a - b).abs().max()) for a, b in zip(ac.parameters(), ac1.parameters()))
            moved = max(float((a - b).abs().max()) for a, b in zip(ac1.parameters(), fresh(arch)[0].parameters()))
            ok = ok and err < 1e-10 and moved > 1e-6
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_data_parallel_ppo_update_equals_the_gathered_update():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
