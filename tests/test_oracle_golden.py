"""CPU: the oracle's env logic against golden vectors recorded from the REFERENCE's own Python code.

tests/golden/*.npz were written by tools/gen_golden.py, which runs mocca_envs' Walker3DStepperEnv / MikeStepperEnv
unmodified (PyBullet and gym stubbed; physics replaced by a tape of post-physics states) inside the reference's
worker wrappers (TimeLimit -> TimeLimitMask -> Monitor, auto-reset).  Here the same tapes go through the oracle
(`sso_step_injected`): every observation, reward, done / bad_transition flag, Monitor episode record, stone index,
dwell counter and terrain table must come out the same.  Tolerances: the reference assembles the robot state in
float32 (robots.py:305-308, 342-345) and returns float32 observations, the oracle computes in float64 and casts
once -> 2e-6 relative on observations; rewards are float64 sums of float32-rounded terms -> 1e-5 absolute.
"""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def _oracle_for(g):
    from oracle.binding import Oracle
    from steppingstone_b200 import model, seeding
    kind, seed = int(g["kind"]), int(g["seed"])
    prm = {"stop_frames": int(g["stop_frames"])}
    if int(g["time_limit"]) == 0:
        prm["max_episode_steps"] = 1 << 30
    M = model.make_model(kind); P = model.make_params(kind, prm)
    o = Oracle(M, P, 1, seeding.env_streams(seed, 0, 1), seeding.aux_streams(kind, seed, 0, 1), threads=1)
    if int(g["curriculum"]) >= 0:
        o.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(int(g["curriculum"]))))
    return o


def seeding_restore(kind):
    from steppingstone_b200 import seeding
    return seeding.temp_state_restore(kind)


def test_golden_files_present():
    assert len(GOLDEN) >= 13


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_env_logic_matches_reference(path):
    g = np.load(path)
    o = _oracle_for(g)
    kind = int(g["kind"])
    obs = o.reset()[0]
    cols = np.ones(60, bool)
    scale = np.ones(60)
    if kind == 1:
        scale[6:9] = 5.7e4          # Mike's locked abdomen: obs = 2(q-mid)/(hi-lo) with hi-lo = 3.5e-5 rad, in float32
    def close(a, b, what, t):
        err = np.abs(a.astype(np.float64) - b.astype(np.float64)) / (scale * np.maximum(1.0, np.abs(b)))
        assert err.max() < 2e-6, "%s differs at step %d: col %d ref %r oracle %r" % (what, t, int(err.argmax()), b[int(err.argmax())], a[int(err.argmax())])
    close(obs, g["obs0"], "reset obs", -1)
    terr_at = list(g["terrain_at"]); terr = g["terrain"]
    st = o.get_state()
    assert np.abs(st["terrain"][0] - terr[0]).max() < 1e-12, "terrain after the first reset"
    T = len(g["rew"])
    n_ep = 0
    ts_step = list(g["ts_step"]) if "ts_step" in g.files else []
    sp_step = list(g["sp_step"]) if "sp_step" in g.files else []
    for t in range(T):
        ob, rew, done, to, er, el = o.step_injected(g["actions"][t][None], g["q"][t][None], g["qd"][t][None], g["contact"][t][None])
        assert bool(done[0]) == bool(g["done"][t]), "done differs at step %d" % t
        assert bool(to[0]) == bool(g["bad"][t]), "bad_transition differs at step %d" % t
        assert abs(float(rew[0]) - g["rew"][t]) < 1e-5 * max(1.0, abs(g["rew"][t])), "reward differs at step %d: ref %r oracle %r" % (t, g["rew"][t], rew[0])
        close(ob[0], g["obs"][t], "obs", t)
        if t in sp_step:
            # update_sample_prob (mike.py:827-830): new table, immediate re-sample iff this step advanced a stone
            from steppingstone_b200 import seeding
            o.set_sample_cdf(seeding.choice_cdf(g["sp_tables"][sp_step.index(t)].astype(np.float64)), apply_now=True)
        if t in ts_step:
            # value-guided sampling probes (create_temp_states_1/_2/create_temp_states, mike.py:699-753): the
            # reference leaves walk_target at the last grid cell, so the NEXT step's reward checks that too
            k = ts_step.index(t); variant = int(g["ts_variant"][k])
            ts = o.create_temp_states(variant, g["ts_grid"], seeding_restore(kind))[0]
            ref = g["ts_%d" % k]
            if variant == 5:
                assert np.abs(ts.astype(np.float64).sum(axis=0) - g["ts_sum_%d" % k]).max() < 2e-6 * 161051 * 5
                ts = ts[::1009]
            assert ts.shape == ref.shape
            err = np.abs(ts.astype(np.float64) - ref) / (scale * np.maximum(1.0, np.abs(ref)))
            assert err.max() < 2e-6, "temp states (variant %d) differ at step %d: %r" % (variant, t, err.max())
        if done[0]:
            n_ep += 1
            assert int(el[0]) == int(g["ep_l"][t])
            assert abs(float(er[0]) - g["ep_r"][t]) < 2e-5 * max(1.0, abs(g["ep_r"][t]))
        if t in terr_at:
            st = o.get_state()
            k = terr_at.index(t)
            assert np.abs(st["terrain"][0] - terr[k]).max() < 1e-12, "terrain_info differs after step %d" % t
            if not done[0]:
                assert int(st["ints"][0, 0]) == int(g["next_step"][t])
                assert int(st["ints"][0, 1]) == int(g["reached_count"][t])
                assert abs(st["scalars"][0, 0] - g["potential"][t]) < 1e-9 * max(1.0, abs(g["potential"][t]))
    assert n_ep == int(g["done"].sum())


def test_seeding_matches_numpy_randomstate():
    """The MT19937 restatement (oracle) reproduces numpy's RandomState.random_sample bit for bit."""
    from oracle.binding import mt_doubles
    from steppingstone_b200 import seeding
    for seed in (0, 8, 123456789):
        rng = seeding.gym_np_random(seed)
        words = seeding.mt_words(rng)
        ref = rng.random_sample(2000)
        out, _ = mt_doubles(words, 2000)
        assert np.array_equal(ref, out)


def test_mirror_function_matches_reference_semantics():
    """get_mirror_function: negate-then-swap as common/envs_utils.py:687-740 (restated with explicit loops)."""
    import torch
    from steppingstone_b200 import mirror
    idx = mirror.get_mirror_indices()
    neg_o, r_o, l_o, neg_a, r_a, l_a = idx
    assert list(neg_o) == [2, 4, 6, 6, 8, 27, 29, 50, 53, 55, 58]          # walker3d_envs.py:1384-1399
    assert list(r_o) == [9, 10, 11, 12, 13, 19, 20, 21, 22, 30, 31, 32, 33, 34, 40, 41, 42, 43, 48]
    assert list(l_o) == [14, 15, 16, 17, 18, 23, 24, 25, 26, 35, 36, 37, 38, 39, 44, 45, 46, 47, 49]
    f = mirror.get_mirror_function(idx)
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(7, 60, generator=g); act = torch.randn(7, 21, generator=g)
    z = torch.zeros(7, 1)
    out = f((obs, z, act, z, z, z, z, z))
    o2 = obs.clone(); o2[:, np.unique(neg_o)] *= -1
    o3 = o2.clone(); o3[:, np.concatenate((r_o, l_o))] = o2[:, np.concatenate((l_o, r_o))]
    a2 = act.clone(); a2[:, neg_a] *= -1
    a3 = a2.clone(); a3[:, np.concatenate((r_a, l_a))] = a2[:, np.concatenate((l_a, r_a))]
    assert torch.equal(out[0], torch.cat([obs, o3])) and torch.equal(out[2], torch.cat([act, a3]))
    assert out[1].shape == (14, 1)
