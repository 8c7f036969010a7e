"""Multi-GPU layout: the N envs are independent (one process per env in the reference, common/envs_utils.py:516-538),
so they shard by contiguous global index with NO collective on the data path.  Rank r owns envs
[first, first+n_local) and seeds them ``seed + global_index`` exactly as ``make_env_fns`` does (envs_utils.py:28).
Collectives appear only at the rollout -> PPO-update boundary (playground/train.py:508-518):
``gather_rollout`` (the north-star plan: NCCL all-gather of the rollout tensors), or the cheaper data-parallel
alternative ``global_advantage_stats`` + gradient all-reduce.  Everything here works with any torch.distributed
backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(rank, world, n_total):
    """Contiguous block of global env indices owned by ``rank``: (first_env, n_local)."""
    base, rem = divmod(int(n_total), int(world))
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def make_sharded_vec_envs(env_id, seed, n_total, rank=None, world=None, **kwargs):
    """This rank's shard of ``make_vec_envs(env_id, seed, n_total, ...)``; global seeds are preserved."""
    from .vec_env import BatchedStepperEnv
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    first, n_local = shard_bounds(rank, world, n_total)
    return BatchedStepperEnv(env_id, seed, n_local, first_env=first, **kwargs)


ROLLOUT_FIELDS = ("observations", "states", "rewards", "value_preds", "returns", "action_log_probs", "actions", "masks", "bad_masks")


def gather_rollout(rollouts, group=None, fields=ROLLOUT_FIELDS, rank_major=False):
    """All-gather the [T(+1), N_local, ...] rollout tensors over the ranks.
    Shards must be equal-sized (N_total divisible by the world size), as in BASELINE config 4 (8 x 8192).
    One ``all_gather_into_tensor`` per field into a [world, T, N_local, ...] buffer (NCCL writes every peer's shard
    straight into place).  Default: one more strided copy puts the env axis second -> dict of [T(+1), N_total, ...], the
    shapes a single-process RolloutStorage of N_total envs would have.  ``rank_major=True`` returns the gathered buffers as
    they are, [world, T(+1), N_local, ...]: the PPO update only ever flattens (time, env) into one sample axis and
    permutes it (storage.py:84-103), so it can consume them without that copy (half the time, no second 32 GB buffer at
    T = 2048)."""
    world = dist.get_world_size(group)
    out = {}
    for name in fields:
        t = getattr(rollouts, name).contiguous()
        buf = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(buf, t, group=group)
        else:                                         # gloo (CPU tests): list form
            dist.all_gather(list(buf.unbind(0)), t, group=group)
        out[name] = buf if rank_major else buf.transpose(0, 1).reshape(t.shape[0], world * t.shape[1], *t.shape[2:])
    return out


def global_advantage_stats(advantages, group=None):
    """mean / std of the advantages over ALL ranks, as ``ppo.py:41-42`` computes them over the whole [T, N] batch
    (unbiased std, torch default)."""
    a = advantages.double()
    s = torch.stack([a.sum(), (a * a).sum(), torch.tensor(float(a.numel()), dtype=torch.float64, device=a.device)])
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean.to(advantages.dtype), var.clamp_min(0).sqrt().to(advantages.dtype)


def reduce_episode_stats(ep_sum, ep_count, group=None):
    """Global mean finished-episode return (drives the curriculum trigger, train.py:503)."""
    s = torch.stack([ep_sum.double().reshape(()), ep_count.double().reshape(())])
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return (s[0] / s[1].clamp_min(1)).item(), int(s[1].item())


def broadcast_table(table, src=0, group=None):
    """Curriculum / sampling tables are decided on rank 0 and broadcast (train.py:118, 269-271)."""
    dist.broadcast(table, src=src, group=group)
    return table


def broadcast_parameters(module, src=0, group=None):
    """Every rank starts from rank ``src``'s policy (the reference has ONE actor_critic, train.py:150-176)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def allreduce_gradients(parameters, group=None):
    """Average the gradients over the ranks with ONE all-reduce of a flat buffer (~0.5 M parameters = 2 MB for the
    reference's 256-wide actor + critic): with equal shards and per-rank mean losses this is the gradient of the mean
    loss over the global minibatch, i.e. what ``algorithms/ppo.py:60-104`` computes on the gathered rollout."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
