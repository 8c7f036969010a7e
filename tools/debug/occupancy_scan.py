"""GPU box: time per ROUND of env_step_kernel as a function of resident warps per SM (is the kernel latency bound -- a round
costs the same with more warps -- or issue bound?).  usage: SS_B200_LIB=variants/libss__DSS_WPB_14.so SS_SMEM_PAD=200000 python tools/debug/occupancy_scan.py 14"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from steppingstone_b200 import BatchedStepperEnv
wpb = int(sys.argv[1]); rounds = 3
N = 148 * wpb * rounds
env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, N, output="torch", terrain_mode=2)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
acts = torch.clamp(torch.randn(32, N, 21, device="cuda", generator=gen) * 0.22, -1, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(40):
    env.step(acts[i % 32])
ms = []
for i in range(64):
    flush.fill_(i & 255)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(acts[(40 + i) % 32]); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
t = float(np.mean(ms))
print("warps/SM %2d  N %5d (%d full rounds)  %.4f ms/step  %.4f ms/round  %.2f M env-steps/s  (%.1f us per env-step-slot)" % (wpb, N, rounds, t, t / rounds, N / t / 1e3, t / rounds * 1e3))
