"""Behavioural pin of the rigid-body arithmetic ("Fixtures that exist"): the policies the reference SHIPS
(playground/models/*.pt, trained in the real PyBullet env) must traverse the stepping stones in this env.

The reference's own success bar is an episode return above 1000 (the curriculum trigger, playground/train.py:503).  A
policy of this kind is a closed-loop controller tuned to one simulator: it falls within a few stones when the contact
model, the collision set or the actuation differ (the negative controls below show that), so walking ~30 stones per
1000-step episode -- 33 is the ceiling, a stone needs 30 dwell frames (walker3d_envs.py:984-990) -- is evidence that
the restated Bullet arithmetic is close to the real one.  It is a behavioural check, not a per-substep one: DESIGN.md 5.

Fixtures: tests/golden/policies/{w3d_latest,w3d_base,mike_latest}.npz -- the actors' weights, read from the checkpoints as DATA by
tools/read_legacy_checkpoint.py (an allow-list pickle walker; nothing the files name is imported or executed).
Forward pass: common/controller.py SoftsignActor.forward (:193-201), deterministic action = the mean (Policy.act
with deterministic=True, :104-110), exactly what playground/enjoy.py runs.
"""
import os

import numpy as np
import pytest

from conftest import make_gpu_env, make_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policies")

# (fixture, robot kind, terrain_mode, curriculum level or None)
CASES = {
    "w3d_latest_shipped_terrain": ("w3d_latest.npz", 0, None, None),        # +-30 deg up/down, walker3d_envs.py:723-730
    "w3d_base_flat_stage": ("w3d_base.npz", 0, 2, None),             # BASELINE configs[1]: the flat-terrain stage
    "mike_latest_curriculum5": ("mike_latest.npz", 1, None, 5),             # BASELINE configs[2]: yaw +-20, pitch +-50
}
RETURN_BAR = 2500.0     # reference bar: 1000 (train.py:503); measured here: 3350-4300
STONES_BAR = 20.0       # measured: 28.9-33.0 stones in the first episode


def load_actor(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: z[k] for k in z.files if k.endswith(("_w", "_b"))}


def actor_numpy(P, obs):
    def softsign(x):
        return x / (1.0 + np.abs(x))
    x = softsign(obs @ P["fc1_w"].T + P["fc1_b"])
    x = softsign(x @ P["fc2_w"].T + P["fc2_b"])
    x = softsign(x @ P["fc3_w"].T + P["fc3_b"])
    x = np.maximum(x @ P["fc4_w"].T + P["fc4_b"], 0.0)
    x = np.maximum(x @ P["fc5_w"].T + P["fc5_b"], 0.0)
    return np.tanh(x @ P["out_w"].T + P["out_b"]).astype(np.float32)


def first_episodes_oracle(o, P, steps, curriculum=None):
    """Deterministic policy rollout on the CPU oracle -> (return, length, stone index) of each env's first episode;
    envs still running after ``steps`` report their running totals."""
    from steppingstone_b200 import seeding
    if curriculum is not None:
        o.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(curriculum)))
    obs = o.reset()
    n = o.n
    ret, length, stones = np.zeros(n), np.zeros(n, int), np.zeros(n, int)
    open_ = np.ones(n, bool)
    run = np.zeros(n)
    for t in range(steps):
        idx_before = o.get_state()["ints"][:, 0].copy()
        obs, rew, done, timeout, ep_ret, ep_len = o.step(actor_numpy(P, obs))
        run += np.where(open_, rew, 0.0)
        fin = open_ & (done != 0)
        ret[fin], length[fin], stones[fin] = ep_ret[fin], ep_len[fin], idx_before[fin]
        open_ &= ~fin
        if not open_.any():
            break
    if open_.any():
        ret[open_], length[open_] = run[open_], steps
        stones[open_] = o.get_state()["ints"][open_, 0]
    return ret, length, stones


@pytest.mark.parametrize("case", sorted(CASES))
def test_shipped_policy_walks_the_stones_in_the_oracle(case):
    name, kind, terrain_mode, curriculum = CASES[case]
    o = make_oracle(kind, 8, terrain_mode=terrain_mode)
    ret, length, stones = first_episodes_oracle(o, load_actor(name), 1000, curriculum)
    o.close()
    assert ret.mean() > RETURN_BAR, (ret, length, stones)
    assert stones.mean() > STONES_BAR, stones
    assert (ret > 1000.0).all(), ret                     # every env clears the reference's own bar


@pytest.mark.parametrize("override", [{"collision_mode": 0}, {"contact_damping": 100.0}], ids=["no_self_collision", "soft_damping_x0.1"])
def test_policy_is_a_sensitive_probe(override):
    """Negative controls: the same policy falls within a few stones when the robot does not collide with itself
    (the reference loads it with URDF_USE_SELF_COLLISION, robots.py:496-500) or when the stones' contact damping
    (bullet_objects.py:67) is ten times smaller -- so the pass above says something about the physics."""
    o = make_oracle(0, 8, params=override)
    ret, length, stones = first_episodes_oracle(o, load_actor("w3d_latest.npz"), 600)
    o.close()
    assert ret.mean() < 1000.0 and stones.mean() < 10.0, (ret, stones)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_shipped_policy_walks_the_stones_on_the_gpu(case):
    """The same rollout through the CUDA path (C-ABI, device tensors), 512 envs, policy in torch on the device."""
    import torch
    name, kind, terrain_mode, curriculum = CASES[case]
    n = 512
    env = make_gpu_env(kind, n, output="torch", terrain_mode=terrain_mode)
    if curriculum is not None:
        env.update_curriculum(curriculum)
    P = {k: torch.tensor(v, device=env.device) for k, v in load_actor(name).items()}

    def actor(x):
        f = torch.nn.functional
        x = f.softsign(f.linear(x, P["fc1_w"], P["fc1_b"]))
        x = f.softsign(f.linear(x, P["fc2_w"], P["fc2_b"]))
        x = f.softsign(f.linear(x, P["fc3_w"], P["fc3_b"]))
        x = torch.relu(f.linear(x, P["fc4_w"], P["fc4_b"]))
        x = torch.relu(f.linear(x, P["fc5_w"], P["fc5_b"]))
        return torch.tanh(f.linear(x, P["out_w"], P["out_b"])).contiguous()

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        obs = env.reset()
        ret = torch.zeros(n, device=env.device); length = torch.zeros(n, dtype=torch.int32, device=env.device)
        open_ = torch.ones(n, dtype=torch.bool, device=env.device)
        for t in range(1000):
            obs, rew, done, info = env.step(actor(obs))
            fin = open_ & (done != 0)
            ret = torch.where(fin, info.episode_r, ret); length = torch.where(fin, info.episode_l, length)
            open_ = open_ & ~fin
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    finished = ~open_.cpu().numpy()
    ret = ret.cpu().numpy(); length = length.cpu().numpy()
    env.close()
    # every first episode ends within 1000 steps (TimeLimit); 512 envs: allow a few early falls, none is typical
    assert finished.all()
    assert ret.mean() > RETURN_BAR, (ret.mean(), np.sort(ret)[:8])
    assert (ret > 1000.0).mean() > 0.9, np.sort(ret)[:16]
    assert length.mean() > 800, length.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["w3d_latest_shipped_terrain", "mike_latest_curriculum5"])
def test_per_substep_parity_on_walking_states(case):
    """PER-SUBSTEP GPU <-> oracle parity on the states that matter: snapshots of the shipped policy WALKING the stones
    in the oracle (feet loaded on tilted stones, swing-leg self-contacts, hinges on their limits), instead of the
    random-action tumbling of tests/test_gpu_parity.py.  From each snapshot, four 1/240 s substeps, each teacher-forced
    from the oracle's state, under the policy's action.  Same fp32 tolerances as test_teacher_forced_substeps:
    1e-3 of the env's velocity scale on generalised velocities, 5e-6 on generalised positions."""
    name, kind, terrain_mode, curriculum = CASES[case]
    from steppingstone_b200 import seeding
    n = 32
    P = load_actor(name)
    walker = make_oracle(kind, n, terrain_mode=terrain_mode)
    if curriculum is not None:
        walker.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(curriculum)))
    prm = dict(n_substeps=1)
    o = make_oracle(kind, n, params=prm, terrain_mode=terrain_mode)
    g = make_gpu_env(kind, n, params=prm, terrain_mode=terrain_mode)
    o.reset(); g.reset(); g.enable_substep_trace(True)
    obs = walker.reset()
    worst_v = worst_q = 0.0
    n_contacts = 0
    for t in range(1, 521):
        a = actor_numpy(P, obs)
        if t in (40, 130, 250, 380, 520):
            o.set_state(walker.get_state())
            for s in range(4):
                st = o.get_state(); g.set_state(st)
                n_contacts += int((st["warm"][:, :8] >= 0).sum())
                o.step(a); g.step(a)
                qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
                vscale = np.maximum(1.0, np.abs(qdo[0]).max(axis=1, keepdims=True))
                worst_v = max(worst_v, float(np.max(np.abs(qdo[0] - qdg[0]) / vscale)))
                worst_q = max(worst_q, float(np.max(np.abs(qo[0] - qg[0]))))
        obs = walker.step(a)[0]
    print("%s: walking-state substeps: max dv (relative) %.2e, max dq %.2e, %d cached contacts seen" % (case, worst_v, worst_q, n_contacts))
    walker.close(); o.close(); g.close()
    assert n_contacts > 100, "the snapshots should be in contact with the stones"
    assert worst_v < 1e-3 and worst_q < 5e-6


def test_example_actor_loads_the_shipped_weights():
    """examples/train_ppo.py's SoftsignActor (``--init-actor``) reproduces the numpy forward pass of this file."""
    import importlib.util
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_ppo_actor", os.path.join(root, "examples", "train_ppo.py"))
    tp = importlib.util.module_from_spec(spec); sys.modules["train_ppo_actor"] = tp; spec.loader.exec_module(tp)
    ac = tp.ActorCritic(60, 21, arch="reference")
    ac.actor.load_npz(os.path.join(GOLD, "mike_latest.npz"))
    obs = np.random.RandomState(0).randn(32, 60).astype(np.float32)
    with torch.no_grad():
        a = ac.dist(torch.as_tensor(obs)).mean.numpy()
    assert np.abs(a - actor_numpy(load_actor("mike_latest.npz"), obs)).max() < 1e-5
