"""Value-guided sampling driver (playground/train.py:224-361) and the single-env view its callers expect."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _AC(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.wa = torch.nn.Parameter(torch.randn(60, 21, generator=g) * 0.05)
        self.wv = torch.nn.Parameter(torch.randn(60, 1, generator=g) * 0.3)

    def act(self, obs, states, masks, deterministic=False):
        a = torch.tanh(obs @ self.wa)
        return obs @ self.wv, a, torch.zeros(obs.shape[0], 1, device=obs.device), states

    def get_value(self, obs, states, masks):
        return obs @ self.wv

    def get_ensemble_values(self, obs, states, masks):                     # common/controller.py:126, one ensemble member
        return (obs @ self.wv)


def _ghost_policy_env(kind_id, seed):
    """an env whose stones are reached quickly: dwell of 2 frames instead of 30"""
    return dict(params=dict(stop_frames=2))


def test_single_env_view_has_the_gym_protocol():
    from steppingstone_b200.sampling import SingleEnv
    env = SingleEnv("mocca_envs:MikeStepperEnv-v0", seed=8, params=dict(max_episode_steps=12))
    obs = env.reset()
    assert obs.shape == (60,) and obs.dtype == np.float32 and env._max_episode_steps == 12
    assert len(env.yaw_samples) == 11 and len(env.pitch_samples) == 11
    n_done = 0
    for t in range(30):
        obs, r, d, info = env.step(np.zeros(21, np.float32))
        assert obs.shape == (60,) and isinstance(r, float) and isinstance(d, bool) and isinstance(info, dict)
        assert isinstance(env.update_terrain, bool)
        if d:
            n_done += 1
            assert "episode" in info and info["episode"]["l"] <= 12
            again = env.reset()                                    # train.py:241-242: the caller resets after done
            assert np.array_equal(again, obs)                      # ... and gets the episode that has already started
    assert n_done >= 2
    assert env.create_temp_states().shape == (121, 60)
    env.close()


def test_value_guided_sampling_matches_the_reference_loop():
    """``value_guided_sample_probs`` on a 1-env torch batch against the LITERAL loop of train.py:229-271 on the single-env
    view (numpy in / numpy out, one env, evaluate_counter >= 5), same seed and policy: the same sampling table."""
    from steppingstone_b200 import BatchedStepperEnv
    from steppingstone_b200.sampling import SingleEnv, value_guided_sample_probs
    dev = torch.device("cuda", 0)
    ac = _AC().to(dev)
    prm = dict(stop_frames=1, max_episode_steps=40)    # a standing start reaches its first stone at once; short episodes give one advance each
    # ---- reference loop (threshold sampling, train.py:229-271)
    evaluate_envs = SingleEnv("mocca_envs:MikeStepperEnv-v0", seed=8, params=prm)
    evaluate_envs.update_curriculum(5)
    obs = evaluate_envs.reset()
    yaw_size, pitch_size = evaluate_envs.yaw_samples.shape[0], evaluate_envs.pitch_samples.shape[0]
    total_metric = torch.zeros(1, yaw_size * pitch_size).to(dev)
    evaluate_counter = 0
    for _ in range(3000):
        obs_t = torch.from_numpy(obs).float().unsqueeze(0).to(dev)
        with torch.no_grad():
            _, action, _, _ = ac.act(obs_t, None, None, deterministic=True)
        cpu_actions = action.squeeze().cpu().numpy()
        obs, reward, done, info = evaluate_envs.step(cpu_actions)
        if done:
            obs = evaluate_envs.reset()
        if evaluate_envs.update_terrain:
            evaluate_counter += 1
            temp_states = evaluate_envs.create_temp_states()
            with torch.no_grad():
                temp_states = torch.from_numpy(temp_states).float().to(dev)
                value_samples = ac.get_ensemble_values(temp_states, None, None)
                metric = value_samples.mean(dim=-1).clone().view(1, yaw_size * pitch_size)
                total_metric += metric
        if evaluate_counter >= 5:
            break
    assert evaluate_counter >= 5, "the test policy never advanced a stone"
    total_metric /= (total_metric.abs().max())
    ref = (-10 * (total_metric - 0.85).abs()).softmax(dim=1).view(yaw_size, pitch_size).cpu().numpy().astype(np.float64)
    # ---- device twin
    eval_dev = BatchedStepperEnv("mocca_envs:MikeStepperEnv-v0", 8, 1, output="torch", params=prm)
    eval_dev.update_curriculum(5)
    probs, count, _ = value_guided_sample_probs(eval_dev, ac, n_events=5, mode="threshold", threshold=0.85)
    assert count == 5 and probs.shape == (11, 11) and abs(probs.sum() - 1.0) < 1e-6
    assert np.abs(probs - ref).max() < 1e-6
    # the table goes to the training envs exactly like train.py:266-271 sends it (one copy per env)
    train_envs = BatchedStepperEnv("mocca_envs:MikeStepperEnv-v0", 8, 64)
    train_envs.reset()
    train_envs.update_sample_prob(np.repeat(probs[None], 64, axis=0))
    # a batch of evaluation envs gives a table of the same kind (more grids per step)
    eval_b = BatchedStepperEnv("mocca_envs:MikeStepperEnv-v0", 8, 32, output="torch", params=prm)
    eval_b.update_curriculum(5)
    p2, c2, _ = value_guided_sample_probs(eval_b, ac, n_events=5, mode="adaptive")
    assert c2 >= 5 and p2.shape == (11, 11) and abs(p2.sum() - 1.0) < 1e-6
