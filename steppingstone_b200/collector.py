"""Device-native rollout collector: the twin of playground/train.py:363-469 with no host round trip.

Per step the reference does: policy on GPU -> actions D2H (:371) -> pipes to N workers (:373) -> results H2D ->
Python loops over N infos (:446-456) -> ``rollouts.insert`` (:459-469).  Here actions, observations, rewards,
masks and episode statistics stay in CUDA tensors; finished-episode returns are accumulated on the device and
read back once per rollout (they drive the curriculum trigger, train.py:503).
"""
import torch


class EpisodeStats:
    """Device-side replacement of ``episode_rewards = deque(maxlen=...)`` fed from infos (train.py:446-454)."""

    def __init__(self, device):
        self.sum = torch.zeros((), dtype=torch.float64, device=device)
        self.count = torch.zeros((), dtype=torch.int64, device=device)
        self.len_sum = torch.zeros((), dtype=torch.int64, device=device)

    def update(self, info):
        d = info.done.bool()
        self.sum += torch.where(d, info.episode_r, torch.zeros_like(info.episode_r)).sum(dtype=torch.float64)
        self.len_sum += torch.where(d, info.episode_l, torch.zeros_like(info.episode_l)).sum(dtype=torch.int64)
        self.count += d.sum()

    def mean(self):
        c = int(self.count.item())
        return (float(self.sum.item()) / c if c else float("nan")), c


@torch.no_grad()
def collect_rollout(envs, actor_critic, rollouts, num_steps, stats=None, deterministic=False):
    """Fill ``rollouts`` with ``num_steps`` transitions from ``envs`` (a torch-mode BatchedStepperEnv)."""
    for step in range(num_steps):
        value, action, action_log_prob, states = actor_critic.act(
            rollouts.observations[step], rollouts.states[step], rollouts.masks[step], deterministic=deterministic)
        obs, reward, done, info = envs.step(action)
        masks = (1.0 - done.to(torch.float32)).unsqueeze(1)                 # train.py:456
        bad_masks = (1.0 - info.bad_transition.to(torch.float32)).unsqueeze(1)   # train.py:446-451
        if stats is not None:
            stats.update(info)
        rollouts.insert(obs, states, action, action_log_prob, value, reward.unsqueeze(1), masks, bad_masks)
    return rollouts
