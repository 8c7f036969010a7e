"""A few device-mode steps for an ncu capture.  usage: python tools/ncu_step.py [collision_mode]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import BatchedStepperEnv
N = 8192
mode = int(sys.argv[1]) if len(sys.argv) > 1 else -1
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
acts = torch.clamp(torch.randn(32, N, 21, device="cuda", generator=gen) * 0.22, -1, 1).contiguous()
env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, N, output="torch", terrain_mode=2,
                        params=None if mode < 0 else dict(collision_mode=mode))
env.reset()
for i in range(24):
    env.step(acts[i])
torch.cuda.synchronize()
