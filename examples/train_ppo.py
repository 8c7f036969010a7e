#!/usr/bin/env python
"""PPO on the device-native path: the caller side of the drop-in (BASELINE config 5 in miniature).

Everything the reference's playground/train.py does per iteration -- act, step N envs, store, GAE, clipped-surrogate
update -- with the env batch, the rollout storage and the GAE scan of this repo and NO host round trip per step
(steppingstone_b200.collector / storage).  The policy and the update restate common/controller.py (DiagGaussian with
log-std -1.5, 256-wide MLPs) and algorithms/ppo.py:40-108 (hyper-parameters of train.py:59-87) in a few lines of torch:
they are callers of the hot path, not part of it.

    python examples/train_ppo.py --env mocca_envs:Walker3DStepperEnv-v0 --envs 4096 --steps 16 --iters 150

Prints one line per iteration: finished-episode mean return / length (Monitor statistics) and env-steps/s.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import make_vec_envs                      # noqa: E402
from steppingstone_b200.collector import EpisodeStats, collect_rollout   # noqa: E402
from steppingstone_b200.storage import RolloutStorage             # noqa: E402


def mlp(sizes, act, gain):
    layers = []
    for i in range(len(sizes) - 1):
        lin = nn.Linear(sizes[i], sizes[i + 1])
        nn.init.orthogonal_(lin.weight, gain=gain if i < len(sizes) - 2 else 1.0)
        nn.init.zeros_(lin.bias)
        layers += [lin] + ([act()] if i < len(sizes) - 2 else [])
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    """Policy.act / get_value / evaluate_actions (common/controller.py:106-150)."""

    def __init__(self, obs_dim, act_dim, h=256):
        super().__init__()
        self.actor = mlp([obs_dim, h, h, h, act_dim], nn.ReLU, nn.init.calculate_gain("relu"))
        self.critic = mlp([obs_dim, h, h, h, h, 1], nn.ReLU, nn.init.calculate_gain("relu"))
        self.logstd = nn.Parameter(torch.full((act_dim,), -1.5))          # DiagGaussian(noise=-1.5), controller.py:21

    def dist(self, obs):
        return torch.distributions.Normal(torch.tanh(self.actor(obs)), self.logstd.exp())

    def act(self, obs, states, masks, deterministic=False):
        d = self.dist(obs)
        a = d.mean if deterministic else d.sample()
        return self.critic(obs), a, d.log_prob(a).sum(-1, keepdim=True), states

    def get_value(self, obs, states, masks):
        return self.critic(obs)

    def evaluate_actions(self, obs, states, masks, action):
        d = self.dist(obs)
        return self.critic(obs), d.log_prob(action).sum(-1, keepdim=True), d.entropy().sum(-1).mean(), states


def ppo_update(ac, opt, rollouts, epochs, num_mini_batch, clip=0.2, max_grad_norm=2.0):
    """algorithms/ppo.py:40-108 (use_clipped_value_loss=False, entropy_coef 0, value_loss_coef 1)."""
    adv = rollouts.returns[:-1] - rollouts.value_preds[:-1]
    adv = (adv - adv.mean()) / (adv.std() + 1e-5)
    vl = al = 0.0
    for _ in range(epochs):
        for obs, states, actions, _, returns, masks, old_logp, adv_t in rollouts.feed_forward_generator(adv, num_mini_batch):
            values, logp, _, _ = ac.evaluate_actions(obs, states, masks, actions)
            ratio = torch.exp(logp - old_logp)
            action_loss = -torch.min(ratio * adv_t, torch.clamp(ratio, 1 - clip, 1 + clip) * adv_t).mean()
            value_loss = 0.5 * (returns - values).pow(2).mean()
            opt.zero_grad(set_to_none=True)
            (value_loss + action_loss).backward()
            nn.utils.clip_grad_norm_(ac.parameters(), max_grad_norm)
            opt.step()
            vl += value_loss.detach(); al += action_loss.detach()
    n = epochs * num_mini_batch
    return float(vl) / n, float(al) / n


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
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    dev = torch.device("cuda", 0)
    envs = make_vec_envs(args.env, args.seed, args.envs, output="torch", device=0, terrain_mode=args.terrain_mode)
    if "Mike" in args.env:
        envs.update_curriculum(args.curriculum)
    ac = ActorCritic(60, 21).to(dev)
    opt = torch.optim.Adam(ac.parameters(), lr=args.lr, eps=1e-5)
    rollouts = RolloutStorage(args.steps, args.envs, (60,), 21, 1, device=dev)
    rollouts.observations[0].copy_(envs.reset())
    num_mini_batch = max(1, args.envs * args.steps // args.mini_batch)
    t_start = time.time(); frames = 0
    curriculum = args.curriculum
    for it in range(args.iters):
        lr = max(args.lr * args.lr_decay ** (frames / 40000.0), 3e-5)
        for g in opt.param_groups:
            g["lr"] = lr
        stats = EpisodeStats(dev)
        t0 = time.time()
        collect_rollout(envs, ac, rollouts, args.steps, stats)
        with torch.no_grad():
            next_value = ac.get_value(rollouts.observations[-1], None, None)
        rollouts.compute_returns(next_value, True, 0.99, 0.95)
        torch.cuda.synchronize(); t1 = time.time()
        vl, al = ppo_update(ac, opt, rollouts, args.epochs, num_mini_batch)
        rollouts.after_update()
        torch.cuda.synchronize(); t2 = time.time()
        frames += args.envs * args.steps
        mean_r, n_ep = stats.mean()
        mean_l = float(stats.len_sum.item()) / max(n_ep, 1)
        # stones reached per 1000 env-steps: a step carrying the step bonus (50 exp(-d/0.25), walker3d_envs.py:1001-1006)
        stones = float((rollouts.rewards > 10.0).sum().item()) / (args.envs * args.steps) * 1000.0
        print("iter %4d frames %9d  episodes %6d  mean_return %8.2f  mean_len %6.1f  rew/step %6.3f  stones/1000steps %5.2f  value_loss %8.3f  "
              "curriculum %d  lr %.1e  collect %.0f env-steps/s (incl. policy)  update %.2fs" % (
                  it, frames, n_ep, mean_r, mean_l, float(rollouts.rewards.mean()), stones, vl, curriculum, lr,
                  args.envs * args.steps / (t1 - t0), t2 - t1), flush=True)
        if args.use_curriculum and n_ep > 0 and mean_r > 1000 and curriculum <= 4:      # train.py:503-506
            curriculum += 1
            envs.update_curriculum(curriculum)
        if args.minutes and time.time() - t_start > args.minutes * 60:
            break
    envs.close()


if __name__ == "__main__":
    main()
