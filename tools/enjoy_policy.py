#!/usr/bin/env python
"""Headless twin of playground/enjoy.py on the GPU env: run a shipped reference policy deterministically over a batch
of envs and report first-episode returns / lengths and the policy-in-the-loop throughput.

    python tools/enjoy_policy.py [--envs 8192] [--steps 1000] > profiles/rXX_policy_rollouts.json

Weights: tests/golden/policies/*.npz (tools/read_legacy_checkpoint.py).  Forward pass: common/controller.py:193-201.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [("w3d_latest", "mocca_envs:Walker3DStepperEnv-v0", None, None, "W3D as shipped (+-30 deg up/down)"),
         ("w3d_base", "mocca_envs:Walker3DStepperEnv-v0", 2, None, "W3D flat stage (BASELINE configs[1])"),
         ("mike_latest", "mocca_envs:MikeStepperEnv-v0", None, 5, "Mike curriculum 5 (BASELINE configs[2])"),
         ("mike_latest", "mocca_envs:MikeStepperEnv-v0", None, 0, "Mike curriculum 0")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    import torch
    from steppingstone_b200 import BatchedStepperEnv
    torch.backends.cuda.matmul.allow_tf32 = False
    f = torch.nn.functional
    out = []
    for tag, env_id, terrain_mode, curriculum, label in CASES:
        env = BatchedStepperEnv(env_id, 8, args.envs, device=0, output="torch", terrain_mode=terrain_mode)
        if curriculum is not None:
            env.update_curriculum(curriculum)
        z = np.load(os.path.join(ROOT, "tests", "golden", "policies", tag + ".npz"))
        P = {k: torch.tensor(z[k], device=env.device) for k in z.files if k.endswith(("_w", "_b"))}

        def actor(x):
            x = f.softsign(f.linear(x, P["fc1_w"], P["fc1_b"]))
            x = f.softsign(f.linear(x, P["fc2_w"], P["fc2_b"]))
            x = f.softsign(f.linear(x, P["fc3_w"], P["fc3_b"]))
            x = torch.relu(f.linear(x, P["fc4_w"], P["fc4_b"]))
            x = torch.relu(f.linear(x, P["fc5_w"], P["fc5_b"]))
            return torch.tanh(f.linear(x, P["out_w"], P["out_b"])).contiguous()

        n = args.envs
        obs = env.reset()
        ret = torch.zeros(n, device=env.device); length = torch.zeros(n, dtype=torch.int32, device=env.device)
        open_ = torch.ones(n, dtype=torch.bool, device=env.device)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(args.steps):
            a = actor(obs)
            ev[t][0].record()
            obs, rew, done, info = env.step(a)
            ev[t][1].record()
            fin = open_ & (done != 0)
            ret = torch.where(fin, info.episode_r, ret); length = torch.where(fin, info.episode_l, length)
            open_ = open_ & ~fin
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        fin = (~open_).cpu().numpy(); r = ret.cpu().numpy()[fin]; l = length.cpu().numpy()[fin]
        ms = np.array([a_.elapsed_time(b_) for a_, b_ in ev])          # env.step alone (sort kernels + env_step_kernel), L2 warm
        walk = ms[min(200, args.steps // 2):]
        out.append({"policy": tag, "env": label, "envs": n, "steps": args.steps, "first_episodes_finished": int(fin.sum()),
                    "return_mean": float(r.mean()), "return_median": float(np.median(r)), "return_p05": float(np.percentile(r, 5)),
                    "frac_above_reference_bar_1000": float((r > 1000).mean()), "length_mean": float(l.mean()),
                    "env_steps_per_s_policy_in_loop": n * args.steps / dt,
                    "env_step_ms_on_walking_states": float(walk.mean()), "env_steps_per_s_device_on_walking_states": n / (float(walk.mean()) * 1e-3)})
        env.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
