"""Device-native rollout collector: the twin of playground/train.py:363-469 with no host round trip.

Per step the reference does: policy on GPU -> actions D2H (:371) -> pipes to N workers (:373) -> results H2D ->
Python loops over N infos (:446-456) -> ``rollouts.insert`` (:459-469).  Here actions, observations, rewards,
masks and episode statistics stay in CUDA tensors; finished-episode returns are kept on the device and read back
once per rollout (they drive the curriculum trigger, train.py:503).  ``GraphedRollout`` captures the whole
act -> step -> insert loop of one rollout in a CUDA graph, so an iteration is ONE graph launch.
"""
import torch


class EpisodeStats:
    """Device-side replacement of ``episode_rewards`` fed from infos (train.py:446-454).

    ``window=None``: totals since construction (sum / count / length of the finished episodes).
    ``window=K``: additionally the reference's ``deque(maxlen=K)`` (train.py:194, K = num_processes) -- a ring of the last
    K finished-episode returns in the reference's append order (step-major, env index ascending within a step), so that
    ``mean()`` is ``np.mean(episode_rewards)``, the number the curriculum trigger compares with 1000 (train.py:503).
    No host synchronisation in ``update``; ``mean()`` reads back."""

    def __init__(self, device, window=None):
        self.sum = torch.zeros((), dtype=torch.float64, device=device)
        self.count = torch.zeros((), dtype=torch.int64, device=device)
        self.len_sum = torch.zeros((), dtype=torch.int64, device=device)
        self.window = int(window) if window else None
        if self.window:
            self.ring = torch.zeros(self.window + 1, dtype=torch.float32, device=device)    # last slot: discard bin

    def update(self, info):
        d = info.done.bool()
        self.sum += torch.where(d, info.episode_r, torch.zeros_like(info.episode_r)).sum(dtype=torch.float64)
        self.len_sum += torch.where(d, info.episode_l, torch.zeros_like(info.episode_l)).sum(dtype=torch.int64)
        if self.window:
            # the k-th finished env of this step goes to ring slot (count + k) mod window; the others to the discard bin
            k = torch.cumsum(d.to(torch.int64), 0) - 1
            slot = torch.where(d, (self.count + k) % self.window, torch.full_like(k, self.window))
            # more finished envs in one step than the window holds: only the last `window` of them survive, as in a deque
            slot = torch.where(d & (k < d.sum() - self.window), torch.full_like(k, self.window), slot)
            self.ring.scatter_(0, slot, info.episode_r.to(torch.float32))
        self.count += d.sum()

    def mean(self):
        """(mean return, number of finished episodes it is over)."""
        c = int(self.count.item())
        if self.window:
            m = min(c, self.window)                       # a full ring holds exactly the last `window` returns
            return (float(self.ring[:m].double().sum().item()) / m if m else float("nan")), m
        return (float(self.sum.item()) / c if c else float("nan")), c


@torch.no_grad()
def collect_rollout(envs, actor_critic, rollouts, num_steps, stats=None, deterministic=False, check=True):
    """Fill ``rollouts`` with ``num_steps`` transitions from ``envs`` (a torch-mode BatchedStepperEnv).
    ``check``: one ``envs.check_errors()`` after the rollout (a non-finite action is refused like robots.py:290 would)."""
    for step in range(num_steps):
        value, action, action_log_prob, states = actor_critic.act(
            rollouts.observations[step], rollouts.states[step], rollouts.masks[step], deterministic=deterministic)
        obs, reward, done, info = envs.step(action)
        masks = (1.0 - done.to(torch.float32)).unsqueeze(1)                 # train.py:456
        bad_masks = (1.0 - info.bad_transition.to(torch.float32)).unsqueeze(1)   # train.py:446-451
        if stats is not None:
            stats.update(info)
        rollouts.insert(obs, states, action, action_log_prob, value, reward.unsqueeze(1), masks, bad_masks)
    if check and hasattr(envs, "check_errors"):
        envs.check_errors()
    return rollouts


class GraphedRollout:
    """One rollout (``num_steps`` x act -> ``ss_step`` -> insert, plus the episode statistics) captured in a CUDA graph.

    The loop of ``collect_rollout`` is launch bound around the env kernel (a policy forward is ~20 small kernels, the
    inserts ~10 copies); replaying it as a graph makes an iteration one launch.  Everything inside is stream-ordered:
    ``ss_step`` only enqueues (a memset and three kernels), the storage indices are fixed per step, torch's Philox
    generator is graph-safe.  The policy's parameters are read at replay time, so optimiser steps between replays are
    seen; the env's curriculum / sampling-table calls happen between replays as usual."""

    def __init__(self, envs, actor_critic, rollouts, num_steps, stats=None, deterministic=False, warmup=2):
        assert rollouts.step == 0, "capture at a rollout boundary"
        self.envs, self.rollouts, self.num_steps = envs, rollouts, num_steps
        dev = rollouts.observations.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                        # warm-up outside the capture (allocator, cuBLAS)
            for _ in range(warmup):
                collect_rollout(envs, actor_critic, rollouts, num_steps, stats, deterministic, check=False)
                rollouts.after_update()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            collect_rollout(envs, actor_critic, rollouts, num_steps, stats, deterministic, check=False)
        self.warmup_env_steps = warmup * num_steps          # env steps taken before the first replay (incl. none from the capture)

    def run(self, check=True):
        self.graph.replay()
        if check and hasattr(self.envs, "check_errors"):
            self.envs.check_errors()
        return self.rollouts
