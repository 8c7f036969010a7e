"""GPU box: cProfile of the numpy-mode step (where do the Python-side microseconds go?)"""
import cProfile, pstats, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from steppingstone_b200 import BatchedStepperEnv
N = 8192
env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, N, terrain_mode=2)
env.reset()
rng = np.random.RandomState(0)
acts = np.clip(rng.randn(8, N, 21) * 0.22, -1, 1).astype(np.float32)
for i in range(10):
    env.step(acts[i % 8])
out4 = np.zeros(4); tot = np.zeros(4)
t0 = time.perf_counter()
for i in range(200):
    o, r, d, inf = env.step(acts[i % 8])
    env._lib.ss_host_timing(env._h, out4.ctypes.data); tot += out4
dt = time.perf_counter() - t0
print("step %.1f us; C %.1f us %s; python %.1f us" % (dt / 200 * 1e6, tot.sum() / 200 * 1e6, (tot / 200 * 1e6).round(1), (dt - tot.sum()) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for i in range(200):
    o, r, d, inf = env.step(acts[i % 8])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
