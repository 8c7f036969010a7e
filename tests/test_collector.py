"""The device-native rollout collector against the reference's own loop (playground/train.py:363-469).

CPU: ``EpisodeStats(window=K)`` against ``deque(maxlen=K)`` fed in the reference's order (train.py:194, 446-454).
GPU: ``collect_rollout`` on a torch-mode env against the LITERAL reference loop on a numpy-mode env with the same seeds
and the same policy -- storages must come out equal tensor for tensor (masks :456, bad_masks :446-451, insert order
:459-469), the episode statistics equal to the deque's mean; then the CUDA-graph replay of the same loop.
"""
from collections import deque

import numpy as np
import pytest
import torch

from steppingstone_b200.collector import EpisodeStats


class _Info:
    def __init__(self, done, bad, r, l):
        self.done, self.bad_transition, self.episode_r, self.episode_l = done, bad, r, l


def test_episode_window_is_the_reference_deque():
    rng = np.random.RandomState(0)
    n, K = 12, 5
    dq = deque(maxlen=K); st = EpisodeStats("cpu", window=K); tot = EpisodeStats("cpu")
    assert np.isnan(st.mean()[0])
    all_r = []
    for t in range(40):
        done = rng.rand(n) < (0.9 if t == 7 else 0.15)           # t == 7: more episodes end in one step than the window holds
        r = rng.randn(n).astype(np.float32) * 100
        for i in range(n):                                       # train.py:446-454: env index order within a step
            if done[i]:
                dq.append(float(r[i])); all_r.append(float(r[i]))
        info = _Info(torch.tensor(done.astype(np.uint8)), torch.zeros(n, dtype=torch.uint8), torch.tensor(r),
                     torch.tensor(rng.randint(1, 1000, n).astype(np.int32)))
        st.update(info); tot.update(info)
        if dq:
            m, c = st.mean()
            assert c == len(dq) and abs(m - np.mean(dq)) < 1e-4 * max(1.0, abs(np.mean(dq)))
    assert tot.mean()[1] == len(all_r) and abs(tot.mean()[0] - np.mean(all_r)) < 1e-4


class _Policy(torch.nn.Module):
    """Stand-in for common/controller.py's actor-critic with its ``act`` signature (train.py:366-370)."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w1 = torch.nn.Parameter(torch.randn(60, 64, generator=g) * 0.2)
        self.wa = torch.nn.Parameter(torch.randn(64, 21, generator=g) * 0.3)
        self.wv = torch.nn.Parameter(torch.randn(64, 1, generator=g) * 0.3)

    def act(self, obs, states, masks, deterministic=False):
        h = torch.tanh(obs @ self.w1)
        mean = torch.tanh(h @ self.wa)
        action = mean if deterministic else mean + 0.2 * torch.randn_like(mean)
        logp = -((action - mean) ** 2).sum(-1, keepdim=True)
        return h @ self.wv, action, logp, states


def _reference_loop(envs, ac, rollouts, num_steps, episode_rewards, dev):
    """playground/train.py:363-469, literally (numpy envs; ``update_current_obs`` with num_stack 1 is a copy)."""
    n = envs.num_envs
    for step in range(num_steps):
        with torch.no_grad():
            value, action, action_log_prob, states = ac.act(rollouts.observations[step], rollouts.states[step],
                                                            rollouts.masks[step], deterministic=True)
        cpu_actions = action.squeeze(1).cpu().numpy()
        obs, reward, done, infos = envs.step(cpu_actions)
        reward = torch.from_numpy(np.expand_dims(np.stack(reward), 1)).float()
        bad_masks = np.ones((n, 1))
        for p_index, info in enumerate(infos):
            keys = info.keys()
            if "bad_transition" in keys:
                bad_masks[p_index] = 0.0
            if "episode" in keys:
                episode_rewards.append(info["episode"]["r"])
        masks = torch.FloatTensor([[0.0] if done_ else [1.0] for done_ in done])
        bad_masks = torch.from_numpy(bad_masks)
        current_obs = torch.from_numpy(obs).float()
        rollouts.insert(current_obs.to(dev), states, action, action_log_prob, value, reward.to(dev), masks.to(dev), bad_masks.float().to(dev))


@pytest.mark.gpu
@pytest.mark.parametrize("graphed", [False, True])
def test_collect_rollout_matches_the_reference_loop(graphed):
    from steppingstone_b200 import BatchedStepperEnv
    from steppingstone_b200.collector import GraphedRollout, collect_rollout
    from steppingstone_b200.storage import RolloutStorage
    dev = torch.device("cuda", 0)
    n, T = 96, 24
    prm = dict(max_episode_steps=17)            # short time limit: bad_transition and auto-reset inside every rollout
    e_ref = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, n, params=prm)
    e_dev = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, n, output="torch", params=prm)
    ac = _Policy().to(dev)
    r_ref = RolloutStorage(T, n, (60,), 21, 1, device=dev); r_dev = RolloutStorage(T, n, (60,), 21, 1, device=dev)
    r_ref.observations[0].copy_(torch.from_numpy(e_ref.reset()).to(dev)); r_dev.observations[0].copy_(e_dev.reset())
    dq = deque(maxlen=n); stats = EpisodeStats(dev, window=n)
    runner = None
    for it in range(3):
        _reference_loop(e_ref, ac, r_ref, T, dq, dev)
        if graphed:
            if runner is None:
                # the capture's warm-up rollouts advance the env: do them on a clone so that e_dev stays in step with e_ref
                st = e_dev.get_state()
                runner = GraphedRollout(e_dev, ac, r_dev, T, stats, deterministic=True, warmup=1)
                e_dev.set_state(st); r_dev.observations[0].copy_(r_ref.observations[0]); r_dev.masks[0].fill_(1.0); r_dev.bad_masks[0].fill_(1.0)
                stats.sum.zero_(); stats.count.zero_(); stats.len_sum.zero_(); stats.ring.zero_()
            runner.run()
        else:
            collect_rollout(e_dev, ac, r_dev, T, stats, deterministic=True)
        torch.cuda.synchronize()
        for name in ("observations", "rewards", "value_preds", "action_log_probs", "actions", "masks", "bad_masks"):
            a, b = getattr(r_ref, name), getattr(r_dev, name)
            assert torch.equal(a, b), "%s differs in rollout %d (max |d| %g)" % (name, it, (a - b).abs().max().item())
        assert (r_dev.bad_masks == 0).any() and (r_dev.masks == 0).any()
        m, c = stats.mean()
        assert c == len(dq) and abs(m - np.mean(dq)) < 1e-4 * max(1.0, abs(np.mean(dq)))
        nv = torch.zeros(n, 1, device=dev)
        r_ref.compute_returns(nv, True, 0.99, 0.95); r_dev.compute_returns(nv, True, 0.99, 0.95)
        assert torch.equal(r_ref.returns, r_dev.returns)
        r_ref.after_update(); r_dev.after_update()
    e_dev.check_errors()


@pytest.mark.gpu
def test_non_finite_action_on_the_device_path_is_flagged():
    """robots.py:290 asserts on a non-finite action; the stream-ordered device path steps it as zero and raises at the
    next check_errors() -- the rollout stays finite (one NaN from the policy used to poison the update)."""
    from steppingstone_b200 import BatchedStepperEnv
    n = 64
    a_env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, n, output="torch")
    b_env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, n, output="torch")
    a_env.reset(); b_env.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    a = torch.clamp(torch.randn(n, 21, device="cuda", generator=gen) * 0.3, -1, 1)
    bad = a.clone(); bad[5, 3] = float("nan"); bad[9, 0] = float("inf")
    zeroed = a.clone(); zeroed[5] = 0; zeroed[9] = 0
    o1, r1, d1, _ = a_env.step(bad); o2, r2, d2, _ = b_env.step(zeroed)
    assert torch.isfinite(o1).all() and torch.isfinite(r1).all()
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    b_env.check_errors()
    with pytest.raises(AssertionError):
        a_env.check_errors()
    a_env.check_errors()                       # the flag was cleared by the failing check


@pytest.mark.gpu
def test_mirror_and_minibatching_on_the_device():
    """SURVEY 8f.2 on CUDA tensors: the mirror function (common/envs_utils.py:687-740 restated as one gather + sign) gives
    what it gives on the host, is an involution on observations / actions, and ``feed_forward_generator``'s device
    permutation visits every sample of the rollout exactly once per epoch (storage.py:84-103)."""
    from steppingstone_b200 import mirror
    from steppingstone_b200.storage import RolloutStorage
    dev = torch.device("cuda", 0)
    f = mirror.get_mirror_function(mirror.get_mirror_indices())
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(33, 60, generator=g); act = torch.randn(33, 21, generator=g); z = torch.zeros(33, 1)
    host = f((obs, z, act, z, z, z, z, z))
    devo = f(tuple(t.to(dev) for t in (obs, z, act, z, z, z, z, z)))
    assert all(torch.equal(a, b.cpu()) for a, b in zip(host, devo))
    m_obs, m_act = devo[0][33:], devo[2][33:]
    twice = f((m_obs, z.to(dev), m_act, z.to(dev), z.to(dev), z.to(dev), z.to(dev), z.to(dev)))
    assert torch.equal(twice[0][33:].cpu(), obs) and torch.equal(twice[2][33:].cpu(), act)
    T, N = 7, 50
    rs = RolloutStorage(T, N, (60,), 21, 1, device=dev)
    rs.observations[:-1, :, 0] = torch.arange(T * N, device=dev, dtype=torch.float32).view(T, N)      # tag every sample
    seen = torch.cat([mb[0][:, 0] for mb in rs.feed_forward_generator(torch.zeros(T, N, 1, device=dev), 4)])
    assert seen.numel() == T * N and torch.equal(seen.sort().values, torch.arange(T * N, device=dev, dtype=torch.float32))
