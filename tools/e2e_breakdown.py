"""Where the end-to-end (host buffers) step time goes.  Run on the GPU box: python tools/e2e_breakdown.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import BatchedStepperEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, N, terrain_mode=2)
env.reset()
rng = np.random.RandomState(0)
acts = [np.clip(rng.randn(N, 21) * 0.22, -1, 1).astype(np.float32) for _ in range(8)]
for i in range(10):
    env.step(acts[i % 8])
K = 100
tt = np.zeros(4); t_all = 0.0
out4 = np.zeros(4)
for i in range(K):
    t0 = time.perf_counter()
    env.step(acts[i % 8])
    t_all += time.perf_counter() - t0
    env._lib.ss_host_timing(env._h, out4.ctypes.data)
    tt += out4
print("N=%d step %.1f us: C total %.1f (act copy %.1f, enqueue %.1f, wait %.1f, unpack %.1f), python %.1f" % (
    N, t_all / K * 1e6, tt.sum() / K * 1e6, *(tt / K * 1e6), (t_all - tt.sum()) / K * 1e6))
