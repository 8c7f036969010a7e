"""Device-time per step for different collision modes / params (GPU box).  usage: python tools/mode_bench.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import BatchedStepperEnv
N = 8192
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
acts = torch.clamp(torch.randn(64, N, 21, device="cuda", generator=gen) * 0.22, -1, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, kw in (("mode0", dict(collision_mode=0)), ("mode1", dict(collision_mode=1)), ("mode2", dict(collision_mode=2))):
    env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, N, output="torch", terrain_mode=2, params=kw)
    env.reset()
    for i in range(16): env.step(acts[i])
    ms = []
    for i in range(64):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[(16 + i) % 64]); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    print(name, "ms/step %.4f" % np.mean(ms))
    env.close()
