"""Seed-exact RNG streams for the batched env.

The reference seeds worker ``rank`` with ``env.seed(seed + rank)`` (common/envs_utils.py:28), which is
``gym.utils.seeding.np_random`` (mocca_envs/env_base.py:131-133): a sha512 hash of the decimal seed
string, its first 8 bytes read as little-endian 32-bit words, fed to ``RandomState.seed`` (MT19937
``init_by_array``).  gym (<0.22, third party, not in the reference tree) is restated here.  The global
numpy stream that ``Walker3DStepperEnv.generate_step_placements`` also draws from
(walker3d_envs.py:728) is ``np.random.seed(seed)`` in the parent before the fork
(common/sacred_utils.py:64-71), so every worker starts from the same copy of it.

The device consumes raw MT19937 state: 624 words + position (``SS_MT_WORDS`` = 625).
"""
import hashlib
import struct

import numpy as np

MT_WORDS = 625


def gym_hash_seed(seed, max_bytes=8):
    """gym.utils.seeding.create_seed + hash_seed for an int seed."""
    seed = int(seed) % 2 ** (8 * max_bytes)
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:max_bytes]
    sizeof_int = 4
    digest += b"\0" * (sizeof_int - len(digest) % sizeof_int)
    words = struct.unpack("{}I".format(len(digest) // sizeof_int), digest)
    return sum(2 ** (sizeof_int * 8 * i) * v for i, v in enumerate(words))


def _int_list_from_bigint(bigint):
    if bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def gym_np_random(seed):
    """The RandomState ``EnvBase.seed(seed)`` creates."""
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(gym_hash_seed(seed)))
    return rng


def mt_words(rng):
    """[625] uint32: MT19937 key + position of a RandomState."""
    name, key, pos, _, _ = rng.get_state()
    assert name == "MT19937"
    out = np.empty(MT_WORDS, dtype=np.uint32)
    out[:624] = key
    out[624] = pos
    return out


def rng_from_words(words):
    rng = np.random.RandomState()
    rng.set_state(("MT19937", np.asarray(words[:624], dtype=np.uint32), int(words[624]), 0, 0.0))
    return rng


def env_streams(seed, first_env, n_envs):
    """[n_envs, 625] MT states of envs ``first_env .. first_env+n_envs-1`` (global numbering)."""
    out = np.empty((n_envs, MT_WORDS), dtype=np.uint32)
    for i in range(n_envs):
        out[i] = mt_words(gym_np_random(seed + first_env + i))
    return out


def global_stream(seed):
    """[625] MT state of ``np.random.seed(seed)`` -- the stream every forked worker inherits."""
    return mt_words(np.random.RandomState(int(seed) % 2 ** 32))


ROBOT_STREAM_OFFSET = 0x5EED0000


def aux_streams(robot_kind, seed, first_env, n_envs):
    """[n_envs, 625] second stream of each env (``ss_create`` mt_aux_host).

    Walker3D: every env gets a copy of the global numpy stream (``generate_step_placements`` draws the
    up/down/flat pitch choice from ``np.random.rand()``, walker3d_envs.py:728; forked workers all inherit the
    parent's ``np.random.seed(seed)`` state, common/sacred_utils.py:64-71).
    Mike: the robot's own ``np_random``.  ``EnvBase.__init__`` links it to the RandomState made by
    ``seed(None)`` (env_base.py:26,65) and ``env.seed(seed+rank)`` replaces only the env's (:131-133), so in
    the reference this stream is seeded from OS entropy; here it is ``RandomState(seed + rank + 0x5EED0000)``
    so that runs are reproducible.  It decides the mirrored start pose (robots.py:846)."""
    out = np.empty((n_envs, MT_WORDS), dtype=np.uint32)
    if robot_kind == 0:
        out[:] = global_stream(seed)
    else:
        for i in range(n_envs):
            out[i] = mt_words(np.random.RandomState((int(seed) + first_env + i + ROBOT_STREAM_OFFSET) % 2 ** 32))
    return out


def curriculum_table(level, grid=11, max_curriculum=5):
    """``update_curriculum`` (walker3d_envs.py:1346-1356 / mike.py:856-866): yaw_pitch_r_tilt_prob."""
    c = min(int(level), max_curriculum)
    half = (grid - 1) // 2
    if c >= half:
        c = half
    p = np.zeros((grid,) * 5)
    w = slice(half - c, half + c + 1)
    p[w, w, 0:c * 2 + 1, w, w] = 1.0 / (c * 2 + 1) ** 5
    return p


def specialist_table(level, grid=11):
    """``update_specialist`` (walker3d_envs.py:604-622 / mike.py)."""
    s = min(int(level), 5)
    prev = s - 1
    half = (grid - 1) // 2
    prob = 1.0 if level == 0 else 1.0 / ((s * 2 + 1) ** 4 - (prev * 2 + 1) ** 4) / 3
    w = slice(half - s, half + s + 1)
    pw = slice(half - prev, half + prev + 1)
    r_index = [0]
    if s > 0:
        r_index += [level * 2 - 1, level * 2]
    p = np.zeros((grid,) * 5)
    # numpy advanced indexing with a list in the middle: assign per r index to keep the same cells
    for r in r_index:
        p[w, w, r, w, w] = prob
    for r in r_index:
        p[pw, pw, r, pw, pw] = 0
    return p


def uniform_table(grid=11):
    return np.ones((grid,) * 5) / grid ** 5


def choice_cdf(p):
    """The cumulative table ``RandomState.choice(p=..., replace=False, size=1)`` searches:
    ``sample_next_next_step`` first renormalises p in place (mike.py:664), then choice() builds
    ``cdf = cumsum(p); cdf /= cdf[-1]``."""
    p = np.asarray(p, dtype=np.float64).reshape(-1).copy()
    p /= p.sum()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.ascontiguousarray(cdf)


def fake_sample_grid():
    """The five ``fake_*_samples`` axes create_temp_states iterates over (walker3d_envs.py:543-547 = mike.py:86-90):
    yaw, pitch, r, x_tilt, y_tilt -> [5, 11] float64."""
    deg = np.pi / 180
    return np.array([np.linspace(-20, 20, num=11) * deg, np.linspace(-50, 50, num=11) * deg,
                     np.linspace(0.65, 1.5, num=11), np.linspace(-20, 20, num=11) * deg,
                     np.linspace(-20, 20, num=11) * deg])


def temp_state_restore(robot_kind):
    """What create_temp_states puts the next-next stone back to (pitch, yaw, dr, x_tilt, y_tilt).  Mike: the values its
    sampler drew for that stone, i.e. the row is unchanged (None).  Walker3D never runs the sampler
    (walker3d_envs.py:1083), so its next_next_* are still the reset values (walker3d_envs.py:836-843) and the
    "restore" flattens the stone: pitch 90 deg, yaw 0, dr 0.8, no tilt."""
    return None if robot_kind == 1 else np.array([np.pi / 2, 0.0, 0.8, 0.0, 0.0])
