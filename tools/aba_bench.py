"""Stand-alone forward-dynamics kernel timing (GPU box).  usage: python tools/aba_bench.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import BatchedStepperEnv
env = BatchedStepperEnv("mocca_envs:Walker3DStepperEnv-v0", 8, 64, output="torch")
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in (8192, 131072):
    q = torch.zeros(N, 28, device="cuda"); q[:, 6] = 1.0
    q[:, 7:] = (torch.rand(N, 21, device="cuda", generator=gen) - 0.5)
    qd = torch.randn(N, 27, device="cuda", generator=gen); tau = torch.randn(N, 21, device="cuda", generator=gen) * 20
    for _ in range(5): env.aba_substep(q, qd, tau)
    ms = []
    for i in range(20):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.aba_substep(q, qd, tau); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = float(np.mean(ms)) * 1e-3
    print("N=%d  %.1f us  %.1f GB/s algorithmic (524 B/env-substep)" % (N, t * 1e6, 524 * N / t / 1e9))
