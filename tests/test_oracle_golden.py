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
    from parity_helpers import golden_params
    from steppingstone_b200 import model, seeding
    kind, seed = int(g["kind"]), int(g["seed"])
    M = model.make_model(kind); P = model.make_params(kind, golden_params(g))
    return Oracle(M, P, 1, seeding.env_streams(seed, 0, 1), seeding.aux_streams(kind, seed, 0, 1), threads=1)


def test_golden_files_present():
    assert len(GOLDEN) >= 14


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_env_logic_matches_reference(path):
    """The tape through the oracle (tests/parity_helpers.py::replay_golden; the GPU replays the same tapes through the
    same function in tests/test_gpu_parity.py::test_gpu_replays_reference_goldens)."""
    from parity_helpers import replay_golden
    g = np.load(path)
    assert replay_golden(g, _oracle_for(g)) == len(g["rew"])


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
