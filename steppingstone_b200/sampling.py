"""Value-guided terrain sampling on the device: the twin of the driver loop in playground/train.py:224-361.

The reference evaluates its critic over the (yaw, pitch) grid of possible next-next stones -- ``create_temp_states`` on ONE
evaluation env, five stone advances in a row -- turns the summed values into a sampling table
(``use_threshold_sampling``: softmax(-10 |v / max|v| - threshold|), train.py:262; ``use_adaptive_sampling``:
softmax(-10 v / max|v|), train.py:357) and hands that table to every training env (``envs.update_sample_prob``).
Here the evaluation env is a (small) torch-mode ``BatchedStepperEnv``: policy, env step, probe grid and critic stay on the
device; the host looks at one counter per step, as the reference looks at ``evaluate_envs.update_terrain``.

``SingleEnv`` is the gym.Env-shaped view of one env that the reference's ``dummy_env`` / ``evaluate_envs`` callers expect
(train.py:96, 229-246, 485): unbatched ``reset`` / ``step``, the ``update_terrain`` attribute, the sample grids.
"""
import numpy as np

from .vec_env import ACT_DIM, OBS_DIM, BatchedStepperEnv


class SingleEnv:
    """One stepping-stone env with the single-env protocol of the reference (gym.Env: ``reset()`` -> obs [60],
    ``step(a [21])`` -> (obs, float reward, bool done, info dict); auto-reset is the CALLER's job there
    (train.py:241-242) -- here the env has already reset itself when ``done`` comes back, which is what that caller
    does next anyway, so its ``reset()`` after ``done`` returns the observation of the episode that already started)."""

    def __init__(self, env_id, seed=0, **kwargs):
        self._env = BatchedStepperEnv(env_id, seed, 1, output="numpy", **kwargs)
        self.observation_space, self.action_space = self._env.observation_space, self._env.action_space
        self.yaw_samples, self.pitch_samples = self._env.yaw_samples, self._env.pitch_samples
        self._max_episode_steps = self._env._max_episode_steps
        self._last_obs = None
        self._fresh = False                   # True right after an auto-reset: the next reset() must not start ANOTHER episode

    def seed(self, seed=None):
        raise NotImplementedError("the seed is fixed at construction (env i of make_vec_envs is seeded seed + i, envs_utils.py:28)")

    def reset(self):
        if not self._fresh or self._last_obs is None:
            self._last_obs = self._env.reset()[0]
        self._fresh = False
        return self._last_obs

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, ACT_DIM)
        obs, rew, done, infos = self._env.step(a)
        self._last_obs, self._fresh = obs[0], bool(done[0])
        return obs[0], float(rew[0]), bool(done[0]), dict(infos[0])

    @property
    def update_terrain(self):
        """``env.update_terrain`` (walker3d_envs.py:1040): the last step advanced a stone."""
        return bool(self._env.get_state()["ints"][0, 5])

    def create_temp_states(self):
        """The (yaw, pitch) probe grid train.py:246-257 views as [yaw_size * pitch_size, 60] (create_temp_states_1)."""
        return self._env.create_temp_states(1)[0]

    def update_curriculum(self, c):
        self._env.update_curriculum(c)

    def update_specialist(self, s):
        self._env.update_specialist(s)

    def update_sample_prob(self, p):
        self._env.update_sample_prob(p)

    def get_mirror_indices(self):
        return self._env.get_mirror_indices()

    def set_mirror(self, mirror):
        pass

    def close(self):
        self._env.close()


def sampling_table_from_metric(total_metric, mode="threshold", threshold=0.85, scale=10.0):
    """train.py:259-262 (threshold) / :355-357 (adaptive) on a [1, G] tensor of summed critic values -> [1, G] probabilities."""
    m = total_metric / total_metric.abs().max()
    if mode == "threshold":
        return (-scale * (m - threshold).abs()).softmax(dim=1)
    if mode == "adaptive":
        return (-scale * m).softmax(dim=1)
    raise ValueError("mode must be 'threshold' or 'adaptive'")


def value_guided_sample_probs(eval_envs, actor_critic, n_events=5, mode="threshold", threshold=0.85, scale=10.0,
                              max_steps=20000, obs=None):
    """Run the evaluation loop of train.py:229-271 on a torch-mode env batch and return (probs [yaw, pitch] f64 numpy,
    number of grid evaluations used, last observation).

    Per step: deterministic action, ``eval_envs.step``; every env that advanced a stone contributes its critic values
    over the (yaw, pitch) probe grid (``create_temp_states(1)``: zero rows for the others) to ``total_metric``; the loop
    ends once ``n_events`` grids have been summed (the reference: ``evaluate_counter >= 5`` with one env, so the sum can
    hold a few more than ``n_events`` grids when several envs of a batch advance in the same step)."""
    import torch
    ys, ps = len(eval_envs.yaw_samples), len(eval_envs.pitch_samples)
    if obs is None:
        obs = eval_envs.reset()
    total = torch.zeros(1, ys * ps, device=obs.device)
    count = 0
    with torch.no_grad():
        for _ in range(max_steps):
            _, action, _, _ = actor_critic.act(obs, None, None, deterministic=True)
            obs, _, _, _ = eval_envs.step(action)
            temp = eval_envs.create_temp_states(1)                              # [E, ys*ps, 60]; zeros where no stone advance
            flagged = temp[:, 0, :].abs().sum(-1) > 0
            n = int(flagged.sum().item())                                       # the one host look per step
            if n:
                values = actor_critic.get_value(temp.reshape(-1, OBS_DIM), None, None).reshape(temp.shape[0], ys * ps)
                total += (values * flagged.unsqueeze(1)).sum(0, keepdim=True)
                count += n
            if count >= n_events:
                break
        else:
            raise RuntimeError("no %d stone advances within %d evaluation steps" % (n_events, max_steps))
    probs = sampling_table_from_metric(total, mode, threshold, scale).view(ys, ps)
    return probs.double().cpu().numpy(), count, obs
