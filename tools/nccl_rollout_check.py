"""BASELINE config 4 in miniature, on real GPUs: every rank steps its shard of the env batch with a random policy into a
device-resident RolloutStorage, then the rollout tensors are all-gathered over NCCL (sharding.gather_rollout) as before
a PPO update.  Checks global seed numbering and that every rank ends up with the same gathered tensors; times the gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_rollout_check.py
"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import sharding
from steppingstone_b200.storage import RolloutStorage

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N_PER, T = int(os.environ.get("SS_ENVS", "8192")), int(os.environ.get("SS_T", "64"))
envs = sharding.make_sharded_vec_envs("mocca_envs:Walker3DStepperEnv-v0", 8, N_PER * world, output="torch", device=local)
assert envs.first_env == rank * N_PER and envs.num_envs == N_PER
ro = RolloutStorage(T, N_PER, (60,), 21, 1, device=dev)
ro.observations[0].copy_(envs.reset())
gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
t0 = time.perf_counter()
for t in range(T):
    a = torch.clamp(torch.randn(N_PER, 21, device=dev, generator=gen) * 0.22, -1, 1)
    obs, rew, done, info = envs.step(a)
    m = (1.0 - done.float()).unsqueeze(1); bm = (1.0 - info.bad_transition.float()).unsqueeze(1)
    ro.insert(obs, torch.zeros(N_PER, 1, device=dev), a, torch.zeros(N_PER, 1, device=dev), torch.zeros(N_PER, 1, device=dev), rew.unsqueeze(1), m, bm)
ro.compute_returns(torch.zeros(N_PER, 1, device=dev), True, 0.99, 0.95)
torch.cuda.synchronize(); dist.barrier()
t_roll = time.perf_counter() - t0
g = sharding.gather_rollout(ro); del g                      # first call: communicator set-up
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g = sharding.gather_rollout(ro); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
nbytes = sum(v.numel() * v.element_size() for v in g.values())
assert g["observations"].shape == (T + 1, N_PER * world, 60)
# every rank must hold identical gathered tensors, and its own shard must sit in its own block
chk = torch.stack([v.double().sum() for v in g.values()])
ref = chk.clone(); dist.broadcast(ref, 0)
assert torch.equal(chk, ref), "gathered rollouts differ between ranks"
assert torch.equal(g["rewards"][:, rank * N_PER:(rank + 1) * N_PER], ro.rewards)
mean, std = sharding.global_advantage_stats(ro.returns[:-1] - ro.value_preds[:-1])
if rank == 0:
    print("world %d: %d envs x %d steps per rank in %.2f s; gather_rollout %.1f MB gathered in %.2f ms = %.1f GB/s into each rank; "
          "global advantage mean %.4f std %.4f" % (world, N_PER, T, t_roll, nbytes / 1e6, ms, nbytes / ms / 1e6, float(mean), float(std)))
envs.close()
dist.destroy_process_group()
