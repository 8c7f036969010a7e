"""A small workload for compute-sanitizer (memcheck / racecheck / initcheck): resets, steps with full collision, a stone
advance with the curriculum sampler, create_temp_states, the host-buffer path.
    compute-sanitizer --tool racecheck python tools/sanitize_step.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import BatchedStepperEnv

n = 48
for env_id in ("mocca_envs:MikeStepperEnv-v0", "mocca_envs:Walker3DStepperEnv-v0"):
    env = BatchedStepperEnv(env_id, 8, n)
    if "Mike" in env_id:
        env.update_curriculum(5)
    env.reset()
    rng = np.random.RandomState(0)
    for t in range(6):
        env.step(np.clip(rng.randn(n, 21) * 0.5, -1, 1).astype(np.float32))
    st = env.get_state()
    st["ints"][:, 1] = 29; st["scalars"][:, 14:16] = 1.0          # force a stone advance on the next step
    tgt = st["terrain"][np.arange(n), st["ints"][:, 0]]
    st["q"][:, 0:2] = tgt[:, 0:2]; st["q"][:, 2] = tgt[:, 2] + 0.9
    lo = np.array(env._model.q_lo); hi = np.array(env._model.q_hi)
    st["q"][::2, 7:] = np.clip(st["q"][::2, 7:] + rng.uniform(-1.2, 1.2, (n // 2, 21)), lo, hi)   # self-contacts
    env.set_state(st)
    env.env_post(np.zeros((n, 21), np.float32))
    ts = env.create_temp_states(1)
    for t in range(3):
        obs, rew, done, infos = env.step(np.clip(rng.randn(n, 21) * 0.5, -1, 1).astype(np.float32))
    assert np.isfinite(obs).all() and ts.shape == (n, 121, 60)
    env.close()
print("sanitize workload ok")
