#!/usr/bin/env python
"""PPO on the device-native path: the caller side of the drop-in (BASELINE config 5 in miniature).

Everything the reference's playground/train.py does per iteration -- act, step N envs, store, GAE, clipped-surrogate
update -- with the env batch, the rollout storage and the GAE scan of this repo and NO host round trip per step
(steppingstone_b200.collector / storage).  The policy and the update restate common/controller.py (DiagGaussian with
log-std -1.5, 256-wide MLPs) and algorithms/ppo.py:40-108 (hyper-parameters of train.py:59-87) in a few lines of torch:
they are callers of the hot path, not part of it.

    python examples/train_ppo.py --env mocca_envs:Walker3DStepperEnv-v0 --envs 4096 --steps 16 --iters 150
    torchrun --nproc-per-node 8 examples/train_ppo.py --envs 8192 ...      # data parallel: env shards stay local,
                                                                           # global advantage stats + gradient all-reduce
    python examples/train_ppo.py --init-actor tests/golden/policies/mike_latest.npz --critic-warmup 20 --logstd -2.5 \
        --env mocca_envs:MikeStepperEnv-v0 --curriculum 5                   # fine-tune a policy the reference ships

Prints one line per iteration: finished-episode mean return / length (Monitor statistics) and env-steps/s.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import make_vec_envs, sharding            # noqa: E402
from steppingstone_b200.collector import EpisodeStats, GraphedRollout, collect_rollout   # noqa: E402
from steppingstone_b200.storage import RolloutStorage             # noqa: E402


def mlp(sizes, act, gain, out_gain=1.0):
    layers = []
    for i in range(len(sizes) - 1):
        lin = nn.Linear(sizes[i], sizes[i + 1])
        nn.init.orthogonal_(lin.weight, gain=gain if i < len(sizes) - 2 else out_gain)
        nn.init.zeros_(lin.bias)
        layers += [lin] + ([act()] if i < len(sizes) - 2 else [])
    return nn.Sequential(*layers)


class SoftsignActor(nn.Module):
    """common/controller.py:165-201: fc1-fc3 softsign, fc4-fc5 relu, ``out`` (tanh applied by the caller here) -- the
    architecture of the policies the reference ships, so that their weights load (``--init-actor``)."""

    def __init__(self, obs_dim, act_dim, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = nn.Linear(obs_dim, h), nn.Linear(h, h), nn.Linear(h, h)
        self.fc4, self.fc5, self.out = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, act_dim)

    def forward(self, x):
        f = nn.functional
        x = f.softsign(self.fc1(x)); x = f.softsign(self.fc2(x)); x = f.softsign(self.fc3(x))
        x = f.relu(self.fc4(x)); x = f.relu(self.fc5(x))
        return self.out(x)

    def load_npz(self, path):
        """Weights as extracted by tools/read_legacy_checkpoint.py (tests/golden/policies/*.npz)."""
        import numpy as np
        z = np.load(path)
        with torch.no_grad():
            for name in ("fc1", "fc2", "fc3", "fc4", "fc5", "out"):
                lin = getattr(self, name)
                lin.weight.copy_(torch.as_tensor(z[name + "_w"])); lin.bias.copy_(torch.as_tensor(z[name + "_b"]))


class ActorCritic(nn.Module):
    """Policy.act / get_value / evaluate_actions (common/controller.py:106-150).  ``arch="reference"``: the shipped
    policies' SoftsignActor; ``"small"``: a 3-hidden-layer ReLU actor (faster, the default of the earlier logs)."""

    def __init__(self, obs_dim, act_dim, h=256, arch="small", logstd=-1.5, out_gain=1.0, fixed_logstd=False):
        super().__init__()
        if arch == "reference":
            self.actor = SoftsignActor(obs_dim, act_dim, h)
            if out_gain != 1.0:
                with torch.no_grad():
                    self.actor.out.weight.mul_(out_gain); self.actor.out.bias.zero_()
        else:
            self.actor = mlp([obs_dim, h, h, h, act_dim], nn.ReLU, nn.init.calculate_gain("relu"), out_gain)
        self.critic = mlp([obs_dim, h, h, h, h, 1], nn.ReLU, nn.init.calculate_gain("relu"))
        self.logstd = nn.Parameter(torch.full((act_dim,), float(logstd)), requires_grad=not fixed_logstd)  # DiagGaussian(noise=-1.5), controller.py:21

    def dist(self, obs):
        return torch.distributions.Normal(torch.tanh(self.actor(obs)), self.logstd.exp(), validate_args=False)   # (validation syncs: not capturable)

    def act(self, obs, states, masks, deterministic=False):
        d = self.dist(obs)
        # (d.sample() is torch.normal with a tensor std, which checks std >= 0 on the host: not capturable in a CUDA graph)
        a = d.mean if deterministic else d.mean + d.stddev * torch.randn_like(d.mean)
        return self.critic(obs), a, d.log_prob(a).sum(-1, keepdim=True), states

    def get_value(self, obs, states, masks):
        return self.critic(obs)

    def evaluate_actions(self, obs, states, masks, action):
        d = self.dist(obs)
        return self.critic(obs), d.log_prob(action).sum(-1, keepdim=True), d.entropy().sum(-1).mean(), states


def ppo_update(ac, opt, rollouts, epochs, num_mini_batch, clip=0.2, max_grad_norm=2.0, distributed=False, actor_frozen=False):
    """algorithms/ppo.py:40-108 (use_clipped_value_loss=False, entropy_coef 0, value_loss_coef 1).  ``distributed``:
    every rank holds its own rollout shard (steppingstone_b200/sharding.py): the advantages are normalised with the
    GLOBAL mean / std and the gradients averaged over the ranks -- the update of the gathered batch without moving it."""
    adv = rollouts.returns[:-1] - rollouts.value_preds[:-1]
    if distributed:
        mean, std = sharding.global_advantage_stats(adv)
        adv = (adv - mean) / (std + 1e-5)
    else:
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
    vl = al = kl = 0.0
    for _ in range(epochs):
        for obs, states, actions, _, returns, masks, old_logp, adv_t in rollouts.feed_forward_generator(adv, num_mini_batch):
            values, logp, _, _ = ac.evaluate_actions(obs, states, masks, actions)
            ratio = torch.exp(logp - old_logp)
            action_loss = -torch.min(ratio * adv_t, torch.clamp(ratio, 1 - clip, 1 + clip) * adv_t).mean()
            value_loss = 0.5 * (returns - values).pow(2).mean()
            opt.zero_grad(set_to_none=True)
            (value_loss + (0.0 if actor_frozen else 1.0) * action_loss).backward()
            if distributed:
                sharding.allreduce_gradients(list(ac.parameters()))
            nn.utils.clip_grad_norm_(ac.parameters(), max_grad_norm)
            opt.step()
            vl += value_loss.detach(); al += action_loss.detach(); kl += (old_logp - logp).mean().detach()
    n = epochs * num_mini_batch
    out = torch.stack([vl, al, kl]) / n
    if distributed:                                   # every rank must see the same numbers (the adaptive lr depends on kl)
        torch.distributed.all_reduce(out); out /= torch.distributed.get_world_size()
    return tuple(out.tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="mocca_envs:Walker3DStepperEnv-v0")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--mini-batch", type=int, default=8192)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--seed", type=int, default=8)
    ap.add_argument("--curriculum", type=int, default=0)
    ap.add_argument("--use-curriculum", action="store_true",
                    help="train.py:503-506: raise the terrain curriculum when the mean episode return passes 1000")
    ap.add_argument("--lr-decay", type=float, default=0.99, help="per 40000 frames, floor 3e-5 (train.py:215-216)")
    ap.add_argument("--terrain-mode", type=int, default=None)
    ap.add_argument("--minutes", type=float, default=0.0, help="stop after this many minutes (0 = run all iterations)")
    ap.add_argument("--arch", default="small", choices=["small", "reference"],
                    help="reference = common/controller.py SoftsignActor, the architecture of the shipped policies")
    ap.add_argument("--init-actor", default=None,
                    help="start from a shipped policy's weights (tests/golden/policies/*.npz; implies --arch reference)")
    ap.add_argument("--critic-warmup", type=int, default=0,
                    help="iterations that train the critic only (use with --init-actor: the critic starts random)")
    ap.add_argument("--logstd", type=float, default=-1.5, help="initial log-std of the Gaussian policy (controller.py:21)")
    # large-batch recipe (none of it is in the reference, whose 40 000-sample batches need none of it; all default off)
    ap.add_argument("--reward-scale", type=float, default=1.0, help="rewards are multiplied by this before GAE (returns of ~500 -> O(5) for the critic)")
    ap.add_argument("--kl-target", type=float, default=0.0, help="> 0: adapt the lr to keep the per-update KL near this (x1.5 / /1.5), instead of --lr-decay")
    ap.add_argument("--out-gain", type=float, default=1.0, help="scale of the actor's last layer at initialisation (0.01: near-zero initial actions)")
    ap.add_argument("--fixed-logstd", action="store_true", help="keep the exploration noise at --logstd instead of learning it (it otherwise collapses once the policy has found how to stand)")
    ap.add_argument("--graph", action="store_true", help="collect each rollout by replaying ONE CUDA graph (collector.GraphedRollout)")
    ap.add_argument("--tf32", action="store_true", help="TF32 matmuls in the policy / update (the env is unaffected)")
    ap.add_argument("--gamma", type=float, default=0.99,
                    help="discount (train.py:78: 0.99).  With the large batches of this trainer 0.99 settles on STANDING (the alive bonus of 2 per "
                         "step outweighs the risk of a first step); 0.95-0.97 learns to walk within a minute (profiles/r02m_*)")
    ap.add_argument("--gamma-final", type=float, default=None, help="anneal the discount linearly to this value ...")
    ap.add_argument("--gamma-anneal", type=int, nargs=2, default=(0, 0), metavar=("START", "LEN"), help="... from iteration START over LEN iterations")
    ap.add_argument("--logfile", default=None, help="rank 0 also appends its iteration lines to this file")
    args = ap.parse_args()
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    # one process per GPU under torchrun: --envs is PER GPU (weak scaling, like bench.py); rank r owns the global env
    # indices [r * envs, (r + 1) * envs) with the reference's seed + index numbering (sharding.py)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    distributed = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed + rank)
    if distributed:
        envs = sharding.make_sharded_vec_envs(args.env, args.seed, args.envs * world, rank, world, output="torch",
                                              device=local, terrain_mode=args.terrain_mode)
    else:
        envs = make_vec_envs(args.env, args.seed, args.envs, output="torch", device=local, terrain_mode=args.terrain_mode)
    if "Mike" in args.env:
        envs.update_curriculum(args.curriculum)
    if args.init_actor:
        args.arch = "reference"
    ac = ActorCritic(60, 21, arch=args.arch, logstd=args.logstd, out_gain=args.out_gain, fixed_logstd=args.fixed_logstd).to(dev)
    if args.init_actor:
        ac.actor.load_npz(args.init_actor)
    if distributed:
        sharding.broadcast_parameters(ac)
    opt = torch.optim.Adam([p_ for p_ in ac.parameters() if p_.requires_grad], lr=args.lr, eps=1e-5)
    rollouts = RolloutStorage(args.steps, args.envs, (60,), 21, 1, device=dev)
    rollouts.observations[0].copy_(envs.reset())
    num_mini_batch = max(1, args.envs * args.steps // args.mini_batch)
    t_start = time.time(); frames = 0
    curriculum = args.curriculum
    stats = EpisodeStats(dev)
    runner = GraphedRollout(envs, ac, rollouts, args.steps, stats, warmup=1) if args.graph else None
    if runner is not None:
        frames += args.envs * args.steps * world          # the capture's warm-up rollout (its data is discarded)
    lr = args.lr
    logf = open(args.logfile, "a") if (args.logfile and rank == 0) else None
    reached = {}                                          # first time / frames at which the mean return passed a mark
    for it in range(args.iters):
        if not args.kl_target:
            lr = max(args.lr * args.lr_decay ** (frames / 40000.0), 3e-5)
        for g in opt.param_groups:
            g["lr"] = lr
        stats.sum.zero_(); stats.count.zero_(); stats.len_sum.zero_()
        t0 = time.time()
        if runner is not None:
            runner.run()
        else:
            collect_rollout(envs, ac, rollouts, args.steps, stats)
        stones = float((rollouts.rewards > 10.0).sum().item()) / (args.envs * args.steps) * 1000.0   # steps carrying a step bonus
        rew_mean = float(rollouts.rewards.mean())
        if args.reward_scale != 1.0:
            rollouts.rewards.mul_(args.reward_scale)
        with torch.no_grad():
            next_value = ac.get_value(rollouts.observations[-1], None, None)
        gamma = args.gamma
        if args.gamma_final is not None and args.gamma_anneal[1] > 0:
            gamma += (args.gamma_final - args.gamma) * min(1.0, max(0.0, (it - args.gamma_anneal[0]) / args.gamma_anneal[1]))
        rollouts.compute_returns(next_value, True, gamma, 0.95)
        torch.cuda.synchronize(); t1 = time.time()
        vl, al, kl = ppo_update(ac, opt, rollouts, args.epochs, num_mini_batch, distributed=distributed,
                                actor_frozen=it < args.critic_warmup)
        rollouts.after_update()
        torch.cuda.synchronize(); t2 = time.time()
        if args.kl_target:                                # same kl on every rank (all-reduced) -> same lr everywhere
            if kl > 2.0 * args.kl_target:
                lr = max(lr / 1.5, 1e-5)
            elif kl < 0.5 * args.kl_target:
                lr = min(lr * 1.5, 1e-2)
        frames += args.envs * args.steps * world
        if distributed:                             # global finished-episode statistics drive the curriculum trigger
            mean_r, n_ep = sharding.reduce_episode_stats(stats.sum, stats.count)
            mean_l, _ = sharding.reduce_episode_stats(stats.len_sum, stats.count)
        else:
            mean_r, n_ep = stats.mean()
            mean_l = float(stats.len_sum.item()) / max(n_ep, 1)
        for mark in (1000, 2000, 3000):
            if n_ep > 0 and mean_r > mark and (mark, curriculum) not in reached:
                reached[(mark, curriculum)] = (time.time() - t_start, frames)
        if rank == 0:
            line = ("iter %4d t %6.1fs frames %11d  episodes %6d  mean_return %8.2f  mean_len %6.1f  rew/step %6.3f  stones/1000steps %5.2f  "
                    "value_loss %9.4f  kl %.4f  logstd %.2f  gamma %.3f  curriculum %d  lr %.1e  collect %.0f env-steps/s (incl. policy, %d GPU)  update %.2fs" % (
                        it, time.time() - t_start, frames, n_ep, mean_r, mean_l, rew_mean, stones, vl, kl, float(ac.logstd.mean()), gamma, curriculum, lr,
                        world * args.envs * args.steps / (t1 - t0), world, t2 - t1))
            print(line, flush=True)
            if logf:
                logf.write(line + "\n"); logf.flush()
        if args.use_curriculum and n_ep > 0 and mean_r > 1000 and curriculum <= 4:      # train.py:503-506
            curriculum += 1
            envs.update_curriculum(curriculum)
        stop = bool(args.minutes and time.time() - t_start > args.minutes * 60)
        if distributed:        # rank 0's clock decides for everybody: a rank that left alone would hang the others' collectives
            flag = torch.tensor([1.0 if stop else 0.0], device=dev)
            torch.distributed.broadcast(flag, src=0)
            stop = bool(flag.item())
        if stop:
            break
    if rank == 0:
        summary = "summary: %d GPU(s), %d frames in %.1f s; " % (world, frames, time.time() - t_start) + "; ".join(
            "mean return > %d at curriculum %d after %.1f s / %d frames" % (m, c, t, f) for (m, c), (t, f) in sorted(reached.items(), key=lambda kv: kv[1]))
        print(summary, flush=True)
        if logf:
            logf.write(summary + "\n"); logf.close()
    envs.close()
    if distributed:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
