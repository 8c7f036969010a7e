#!/usr/bin/env python
"""bench.py -- env-steps/s of the stepping-stone hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 256 --warmup 32
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle port on all host threads

A "step" = one control step (4 physics substeps at 240 Hz, 20 solver iterations each, observation, reward,
termination, auto-reset) of every env of the batch.  Workload at N=1 is BASELINE.json configs[1]:
Walker3DStepperEnv-v0, 8192 envs per GPU, flat-terrain stage, actions clip(N(0, e^-1.5), -1, 1).
Prints ONE JSON line on rank 0 (contract in the task statement; keys documented in DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

_RESULT_FD = None


def claim_stdout():
    """Rank 0 prints ONE JSON line on stdout and nothing else may: libraries write there too (NCCL prints its
    "NCCL version ..." banner on stdout at NCCL_DEBUG=VERSION/WARN/INFO).  Keep a private duplicate of the real stdout
    for the result line and point file descriptor 1 -- Python's sys.stdout and every C library alike -- at stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        try:
            sys.stdout.flush()
            fd = os.dup(1)
            os.dup2(2, 1)
            _RESULT_FD = fd
        except OSError:                # no usable stderr: leave stdout as it is rather than fail the run
            _RESULT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    fd = 1 if _RESULT_FD is None else _RESULT_FD
    while data:
        data = data[os.write(fd, data):]


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ENV_STEP = 884      # fused env-step boundary, fp32
ALG_BYTES_PER_ENV_SUBSTEP = 524   # aba_substep kernel boundary, fp32
H2D_PER_ENV = 21 * 4
D2H_PER_ENV = 64 * 4              # one packed result record per env: obs 60 f32, reward, episode return, length, done|timeout
WORKLOADS = {
    "w3d-flat": dict(env_id="mocca_envs:Walker3DStepperEnv-v0", kind=0, terrain_mode=2, curriculum=None,
                     name="Walker3DStepperEnv-v0 flat-terrain stage, 240 Hz / 4 substeps, 20 PGS iters"),
    "w3d": dict(env_id="mocca_envs:Walker3DStepperEnv-v0", kind=0, terrain_mode=None, curriculum=None,
                name="Walker3DStepperEnv-v0 as shipped (+-30 deg up/down terrain)"),
    "mike": dict(env_id="mocca_envs:MikeStepperEnv-v0", kind=1, terrain_mode=None, curriculum=5,
                 name="MikeStepperEnv-v0 use_curriculum (level 5: yaw +-20, pitch +-50)"),
}
# The reference's shipped policies (weights read as data: tests/golden/policies/*.npz, tools/read_legacy_checkpoint.py)
# as the action source: the characters WALK the stones (~6 contact points per env and substep) instead of flailing under
# random actions -- the state mix a training run spends its time in once the policy works.
WALK_POLICY = {"w3d-flat": "w3d_base", "w3d": "w3d_latest", "mike": "mike_latest"}
for _k in list(WALK_POLICY):
    WORKLOADS[_k + "-walk"] = dict(WORKLOADS[_k], policy=WALK_POLICY[_k],
                                   name=WORKLOADS[_k]["name"] + ", actions from the reference's shipped policy " + WALK_POLICY[_k] + " (walking states)")
WALK_WARMUP = 200      # policy steps before the timed region of a walking workload: the gait is established, stones are being reached


def load_actor(tag, dev):
    """common/controller.py:193-201 SoftsignActor.forward, deterministic (action = mean), as playground/enjoy.py runs it."""
    import torch
    f = torch.nn.functional
    z = np.load(os.path.join(ROOT, "tests", "golden", "policies", tag + ".npz"))
    P = {k: torch.tensor(z[k], device=dev) for k in z.files if k.endswith(("_w", "_b"))}

    @torch.no_grad()
    def actor(x):
        x = f.softsign(f.linear(x, P["fc1_w"], P["fc1_b"])); x = f.softsign(f.linear(x, P["fc2_w"], P["fc2_b"]))
        x = f.softsign(f.linear(x, P["fc3_w"], P["fc3_b"]))
        x = torch.relu(f.linear(x, P["fc4_w"], P["fc4_b"])); x = torch.relu(f.linear(x, P["fc5_w"], P["fc5_b"]))
        return torch.tanh(f.linear(x, P["out_w"], P["out_b"])).contiguous()
    return actor


def timed_steps(env, next_action, K, flush, barrier, obs0):
    """K control steps, each bracketed by its own CUDA-event pair on the launching stream; the L2 flush and the action
    source (ring lookup or policy forward) sit OUTSIDE the pair.  Returns per-step ms, wall seconds."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    obs = obs0
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xff)                      # L2 flush between timed steps (outside the event pair)
        a = next_action(i, obs)
        ev[i][0].record()
        obs, _, _, _ = env.step(a)
        ev[i][1].record()
    barrier()
    return np.array([a_.elapsed_time(b_) for a_, b_ in ev]), time.perf_counter() - t0, obs


def exchange_block(env, dev, world, rank, N, T, acts):
    """BASELINE configs[3]: the rollout -> PPO-update boundary at `world` GPUs (playground/train.py:508-518).  Every rank
    fills a device-resident [T+1, N, .] RolloutStorage by stepping its env shard, runs the GAE scan, then the exchange is
    timed both ways: (a) the north star's plan -- NCCL all-gather of the rollout tensors (sharding.gather_rollout: every
    rank ends up with the global [T+1, world*N, .] batch); (b) the data-parallel alternative -- shards stay local, global
    advantage statistics + one flat gradient all-reduce per minibatch (sharding.global_advantage_stats /
    allreduce_gradients).  Device-timed with CUDA events after a warm-up call, max over ranks."""
    import torch
    import torch.distributed as dist
    from steppingstone_b200 import sharding
    from steppingstone_b200.storage import RolloutStorage
    free, _ = torch.cuda.mem_get_info(dev)
    per_rank = (T + 1) * N * (60 + 21 + 7) * 4
    need = per_rank * (2 * world + 1)                 # own shard + gathered copy + the largest field's staging buffer
    asked = T
    while need > 0.8 * free and T > 64:
        T //= 2
        per_rank = (T + 1) * N * (60 + 21 + 7) * 4; need = per_rank * (2 * world + 1)
    ro = RolloutStorage(T, N, (60,), 21, 1, device=dev)
    ro.observations[0].copy_(env.reset())
    z1 = torch.zeros(N, 1, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    for t in range(T):
        a = acts[t % acts.shape[0]]
        obs, rew, done, info = env.step(a)
        ro.insert(obs, z1, a, z1, z1, rew.unsqueeze(1), (1.0 - done.float()).unsqueeze(1), (1.0 - info.bad_transition.float()).unsqueeze(1))
    ro.value_preds.copy_(torch.randn(ro.value_preds.shape, device=dev, generator=gen))
    torch.cuda.synchronize(dev); fill_s = time.perf_counter() - t0

    def timed(fn, reps=1):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize(dev)
        a_.record()
        for _ in range(reps):
            out = fn()
        b_.record(); torch.cuda.synchronize(dev)
        tt = torch.tensor([a_.elapsed_time(b_) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), out

    gae_ms, _ = timed(lambda: ro.compute_returns(z1, True, 0.99, 0.95))
    # (a) all-gather of the rollout (small warm-up first: communicator + allocator)
    sharding.gather_rollout(ro, fields=("rewards",))
    ms_g, g = timed(lambda: sharding.gather_rollout(ro))
    out_bytes = sum(v.numel() * v.element_size() for v in g.values())
    ok = tuple(g["observations"].shape) == (T + 1, world * N, 60) and bool(torch.equal(g["rewards"][:, rank * N:(rank + 1) * N], ro.rewards))
    chk = g["rewards"].double().sum(); ref = chk.clone(); dist.broadcast(ref, 0)
    ok = ok and bool(torch.equal(chk, ref))
    del g
    torch.cuda.empty_cache()
    ms_r, g = timed(lambda: sharding.gather_rollout(ro, rank_major=True))          # no env-axis transpose: [world, T+1, N, .]
    ok = ok and tuple(g["observations"].shape) == (world, T + 1, N, 60) and bool(torch.equal(g["rewards"][rank], ro.rewards))
    del g
    torch.cuda.empty_cache()
    # (b) data-parallel alternative: advantage statistics + gradient all-reduce of the reference's actor + critic (~0.5 M parameters)
    adv = ro.returns[:-1] - ro.value_preds[:-1]
    sharding.global_advantage_stats(adv)
    ms_adv, _ = timed(lambda: sharding.global_advantage_stats(adv), reps=5)
    n_param = (60 * 256 + 256 + 4 * (256 * 256 + 256) + 256 * 21 + 21) + (60 * 256 + 256 + 3 * (256 * 256 + 256) + 256 + 1) + 21
    params = [torch.nn.Parameter(torch.zeros(n_param, device=dev))]
    params[0].grad = torch.randn(n_param, device=dev, generator=gen)
    sharding.allreduce_gradients(params)
    ms_gr, _ = timed(lambda: sharding.allreduce_gradients(params), reps=20)
    n_mb = 10 * max(1, N * T // 32768)               # train.py:59-87: 10 epochs, minibatches of ~32k samples per rank
    return {"rollout_T": T, "rollout_T_asked": asked, "envs_per_rank": N, "ranks": world, "shard_bytes_per_rank": per_rank,
            "fill_s": fill_s, "fill_env_steps_per_s": world * N * T / fill_s, "gae_ms": gae_ms, "verified": ok,
            "gather": {"collective": "ncclAllGather (all_gather_into_tensor per field, 9 fields) + one strided copy per field",
                       "bytes_received_per_rank": out_bytes, "ms": ms_g, "GBps_into_each_rank": out_bytes / ms_g / 1e6,
                       "nvlink5_per_direction_GBps": 900.0, "frac_of_nvlink": out_bytes * (world - 1) / world / ms_g / 1e6 / 900.0,
                       "rank_major_ms": ms_r, "rank_major_GBps_into_each_rank": out_bytes / ms_r / 1e6,
                       "rank_major_frac_of_nvlink": out_bytes * (world - 1) / world / ms_r / 1e6 / 900.0,
                       "rank_major_note": "gathered buffers kept [world, T+1, N, .] (no strided copy): what a PPO update that flattens (time, env) needs"},
            "data_parallel": {"collective": "ncclAllReduce", "advantage_stats_ms": ms_adv, "grad_allreduce_ms": ms_gr,
                              "grad_bytes": n_param * 4, "allreduces_per_update": n_mb, "per_update_ms": ms_adv + n_mb * ms_gr}}


def mean_contacts(env):
    """contact points per env in the last substep (the warm-start list), from the device state."""
    w = env.get_state()["warm"][:, :8]
    return float((w >= 0).sum(axis=1).mean())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of ONE GPU sampled DURING the timed region.  NVML in-process (a query is
    ~0.1 ms, so even the 0.1 s timed region of a short run holds tens of samples); an `nvidia-smi -lms` child
    process is the fallback when the NVML binding is missing (its first sample alone takes ~1 s)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, uuid=None, period_s=0.004):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.rows, self.proc, self.thread = [], None, None       # rows: (sm_mhz, sm_max_mhz, set(reasons))
        self.source, self._nv, self._h, self._mx = None, None, None, None
        self._run = threading.Event()

    # ---- NVML ----
    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            for u in (self.uuid, "GPU-" + self.uuid):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(u if isinstance(u, bytes) else u.encode()); break
                except Exception:
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(u); break
                    except Exception:
                        h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self._mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)   # fail here, not in the thread
        self._nv, self._h = pynvml, h

    def _nvml_sample(self):
        nv, h = self._nv, self._h
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        return sm, self._mx, {n for n, b in self.BITS.items() if mask & b}

    def _nvml_loop(self):
        while self._run.is_set():
            try:
                self.rows.append(self._nvml_sample())
            except Exception:
                pass
            time.sleep(self.period)

    # ---- nvidia-smi fallback ----
    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) < 6:
                continue
            try:
                self.rows.append((float(r[0]), float(r[1]),
                                  {n for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")}))
            except ValueError:
                continue

    def start(self):
        try:
            self._nvml_open()
            self.source = "nvml"
            self._run.set()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self._nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid and ("GPU-" + self.uuid) or str(self.index),
                                          "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
            t_wait = time.perf_counter()
            while not self.rows and time.perf_counter() - t_wait < 8.0:   # its first sample takes ~1 s
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """Forget the samples taken so far (call right before the timed region)."""
        self.rows.clear()

    def stop(self):
        if self._nv is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}
        self._run.clear()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        rows = list(self.rows)
        sm = [r[0] for r in rows]
        reasons = set().union(*[r[2] for r in rows]) if rows else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in rows) if rows else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def cpu_port_rate(wl, n_envs, steps, warmup, threads=0, seed=8):
    """env-steps/s of the CPU oracle (C port of the reference algorithm) on the host cores.  The thread count is
    passed explicitly (OpenMP num_threads clause): torchrun exports OMP_NUM_THREADS=1 to its workers."""
    threads = threads or (os.cpu_count() or 1)
    from oracle.binding import Oracle
    from steppingstone_b200 import model, seeding
    M = model.make_model(wl["kind"])
    P = model.make_params(wl["kind"], None, wl["terrain_mode"])
    o = Oracle(M, P, n_envs, seeding.env_streams(seed, 0, n_envs), seeding.aux_streams(wl["kind"], seed, 0, n_envs), threads=threads)
    if wl["curriculum"] is not None:
        o.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(wl["curriculum"])))
    o.reset()
    rng = np.random.RandomState(0)
    acts = [np.clip(rng.randn(n_envs, 21) * 0.22, -1, 1).astype(np.float32) for _ in range(8)]
    for i in range(warmup):
        o.step(acts[i % 8])
    t0 = time.perf_counter()
    for i in range(steps):
        o.step(acts[i % 8])
    dt = time.perf_counter() - t0
    o.close()
    return n_envs * steps / dt, dt


def _real_reference_worker(env_id, seed, steps, warmup, q):
    """One reference env (the unmodified mocca_envs class on PyBullet) stepped in its own process, as a SubprocVecEnv
    worker does (common/envs_utils.py:623-675): random actions of the workload's distribution, auto-reset on done."""
    import gym
    env = gym.make(env_id); env.seed(seed); env.reset()
    rng = np.random.RandomState(seed)
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        _, _, d, _ = env.step(np.clip(rng.randn(21) * 0.22, -1, 1).astype(np.float32))
        if d:
            env.reset()
    q.put(time.perf_counter() - t0)


def real_reference_rate(wl, steps, warmup):
    """The REAL reference (PyBullet) on all host cores, if it can run here: needs `pybullet`, `gym` (< 0.22) and the
    reference's `mocca_envs` package (baseline/_ref: `pip install --no-deps --target baseline/_ref` of the reference's
    .environments/, done by __graft_entry__.build() where /root/reference exists).  Returns (rate, cores, None) or
    (None, None, reason).  Probed on the GPU box in round 2: profiles/r02a_probe_reference_deps.log -- neither module
    exists there, no wheel in /opt/wheelhouse, no network."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    missing = []
    for mod in ("pybullet", "gym", "mocca_envs"):
        try:
            __import__(mod)
        except Exception as exc:                    # ModuleNotFoundError, or mocca_envs failing on ITS imports
            missing.append("%s: %s" % (mod, repr(exc)[:80]))
    if missing:
        return None, None, "; ".join(missing)
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    q = mp.Queue()
    procs = [mp.Process(target=_real_reference_worker, args=(wl["env_id"], 8 + i, steps, warmup, q)) for i in range(cores)]
    for p_ in procs:
        p_.start()
    times = [q.get() for _ in procs]
    for p_ in procs:
        p_.join()
    return cores * steps / max(times), cores, None


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  First choice: the real
    reference (real_reference_rate).  PyBullet, where the reference's physics lives, is absent from the reference tree,
    from this image and from the GPU box, so in practice this arm times the oracle port (kind "port") on all host threads
    -- and says so, with the reason, on its line.  Same workload and batch size as the GPU arm (`--ref-envs` = --envs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = args.ref_envs or args.envs
    kind, probe = "port", None
    rate, rcores, why = real_reference_rate(wl, args.steps, args.warmup) if "policy" not in wl else (None, None, "walking workloads use the port")
    if rate is not None:
        kind, cores, dt = "reference", rcores, args.steps * rcores / rate
        sample = "%d reference envs (one process per core, gym.make + PyBullet) x %d steps each" % (rcores, args.steps)
        n_envs = rcores
    else:
        probe = {"real_reference": "unavailable", "why": why, "gpu_box_probe": "profiles/r02a_probe_reference_deps.log"}
        rate, dt = cpu_port_rate(wl, n_envs, args.steps, args.warmup)
        sample = "%d envs x %d steps (the GPU arm steps %d envs per GPU), %d OpenMP threads" % (n_envs, args.steps, args.envs, cores)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "envs_per_step": n_envs, "actions": "clip(N(0,0.22),-1,1)"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if probe:
        line["reference_probe"] = probe
    emit(line)


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from steppingstone_b200 import BatchedStepperEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, K, W = args.envs, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    env = BatchedStepperEnv(wl["env_id"], 8, N, device=local, output="torch", first_env=rank * N,
                            terrain_mode=wl["terrain_mode"])
    if wl["curriculum"] is not None:
        env.update_curriculum(wl["curriculum"])
    obs = env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(rank)
    ring = 64
    acts = torch.clamp(torch.randn(ring, N, 21, device=dev, generator=gen) * 0.22, -1, 1).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    actor = load_actor(wl["policy"], dev) if "policy" in wl else None
    if actor is not None:
        W = max(W, WALK_WARMUP)
        next_action = lambda i, o: actor(o)
    else:
        next_action = lambda i, o: acts[(W + i) % ring]

    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)   # CUDA_VISIBLE_DEVICES may renumber the GPUs
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid); sampler.start()
    sampler.mark()
    for i in range(W):
        obs, _, _, _ = env.step(actor(obs) if actor is not None else acts[i % ring])
    barrier()
    l0 = env.launch_count
    if sampler.source == "nvml":
        sampler.mark()                             # NVML: samples of the timed region only (the fallback keeps the warm-up's too)
    step_ms, t_wall, obs = timed_steps(env, next_action, K, flush, barrier, obs)
    launches = env.launch_count - l0
    clocks = sampler.stop()
    contacts_main = mean_contacts(env)
    if world > 1:                                  # every rank samples its own GPU; rank 0 reports the slowest + all reasons
        names = ClockSampler.NAMES
        cmin = torch.tensor([clocks["sm_mhz"] if clocks["sm_mhz"] is not None else float("inf")], dtype=torch.float64, device=dev)
        flags = torch.tensor([1.0 if n_ in clocks["reasons"] else 0.0 for n_ in names], dtype=torch.float64, device=dev)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        other = [r_ for r_ in clocks["reasons"] if r_ not in names]
        clocks = dict(clocks, reasons=sorted(other + [n_ for n_, f_ in zip(names, flags.tolist()) if f_ > 0]),
                      sm_mhz_rank0=clocks["sm_mhz"], ranks=world)
        clocks["sm_mhz"] = None if cmin.item() == float("inf") else float(cmin.item())
    t_dev = float(step_ms.sum()) * 1e-3
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    value = world * N * K / t_max

    # ---- end to end through the reference-facing call with HOST buffers (ss_step_host) ----
    env_h = BatchedStepperEnv(wl["env_id"], 8, N, device=local, output="numpy", first_env=rank * N,
                              terrain_mode=wl["terrain_mode"])
    if wl["curriculum"] is not None:
        env_h.update_curriculum(wl["curriculum"])
    env_h.reset()
    ring_h = 8                                     # host action batches in rotation (5.5 MB: a policy's fresh output is cache-warm too)
    acts_h = acts[:ring_h].cpu().numpy()
    Ke = max(8, min(K, 64))
    for i in range(3):
        obs_h, rew_h, done_h, infos_h = env_h.step(acts_h[i % ring_h])     # held like in the timed loop (output array rotation)
    barrier()
    host4 = np.zeros(4); out4 = np.zeros(4)
    t0 = time.perf_counter()
    for i in range(Ke):
        obs_h, rew_h, done_h, infos_h = env_h.step(acts_h[(3 + i) % ring_h])
        env_h._lib.ss_host_timing(env_h._h, out4.ctypes.data); host4 += out4
    barrier()
    t_e2e = time.perf_counter() - t0
    host_us = dict(zip(("action_check_copy", "enqueue", "wait_device", "unpack"), (host4 / Ke * 1e6).round(1).tolist()))
    host_us["python"] = round((t_e2e - host4.sum()) / Ke * 1e6, 1)
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * Ke / float(te.item())
    env_h.close()

    # ---- stand-alone ABA (forward dynamics) kernel: the 524 B / env-substep boundary ----
    q = torch.zeros(N, 28, device=dev); q[:, 6] = 1.0
    q[:, 7:] = (torch.rand(N, 21, device=dev, generator=gen) - 0.5)
    qd = torch.randn(N, 27, device=dev, generator=gen)
    tau = torch.randn(N, 21, device=dev, generator=gen) * 20
    for _ in range(5):
        env.aba_substep(q, qd, tau)
    torch.cuda.synchronize(dev)
    aba_ms = []
    for i in range(20):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.aba_substep(q, qd, tau); b.record()
        torch.cuda.synchronize(dev)
        aba_ms.append(a.elapsed_time(b))
    aba_s = float(np.mean(aba_ms)) * 1e-3
    # the same kernel on a batch large enough to amortise launch + tail effects (16 x the workload's env count)
    NL = 16 * N
    qL = q.repeat(16, 1).contiguous(); qdL = qd.repeat(16, 1).contiguous(); tauL = tau.repeat(16, 1).contiguous()
    for _ in range(3):
        env.aba_substep(qL, qdL, tauL)
    torch.cuda.synchronize(dev)
    abaL_ms = []
    for i in range(10):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.aba_substep(qL, qdL, tauL); b.record()
        torch.cuda.synchronize(dev)
        abaL_ms.append(a.elapsed_time(b))
    abaL_s = float(np.mean(abaL_ms)) * 1e-3

    # ---- the same kernel on WALKING states (the random-action mix is light): the reference's shipped policy
    #      drives a second env of the same workload; only env.step sits inside the event pairs ----
    walking = None
    wk = args.workload + "-walk"
    if actor is None and wk in WORKLOADS and not args.no_walking:
        try:
            wlw = WORKLOADS[wk]
            env_w = BatchedStepperEnv(wlw["env_id"], 8, N, device=local, output="torch", first_env=rank * N, terrain_mode=wlw["terrain_mode"])
            if wlw["curriculum"] is not None:
                env_w.update_curriculum(wlw["curriculum"])
            actor_w = load_actor(wlw["policy"], dev)
            ob = env_w.reset()
            for _ in range(WALK_WARMUP):
                ob, _, _, _ = env_w.step(actor_w(ob))
            Kw = max(16, min(K, 128))
            ms_w, _, ob = timed_steps(env_w, lambda i, o: actor_w(o), Kw, flush, barrier, ob)
            tw = torch.tensor([float(ms_w.sum()) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            walking = {"workload": wlw["name"], "value": world * N * Kw / float(tw.item()), "unit": "env-steps/s",
                       "ms_per_step": float(tw.item()) / Kw * 1e3, "steps": Kw, "policy_warmup_steps": WALK_WARMUP,
                       "contacts_per_env": mean_contacts(env_w)}
            env_w.close()
        except Exception as exc:
            walking = {"error": repr(exc)[:200]}

    # ---- create_temp_states fill kernel (value-guided sampling probes): the one HBM-bound kernel of the path,
    #      240 B written per (env, grid cell); 48 envs x 161051 cells = 1.86 GB per launch (> L2, no flush needed) ----
    tsf = None
    try:
        env_t = BatchedStepperEnv(WORKLOADS["mike"]["env_id"], 8, 48, device=local, output="torch")
        env_t.update_curriculum(5)
        env_t.reset()
        st = env_t.get_state()
        st["ints"][:, 5] = 1                                   # update_terrain flag: every env produces its grid
        env_t.set_state(st)
        from steppingstone_b200 import native as _nat, seeding as _seed
        cells = 11 ** 5
        out_t = torch.empty((48, cells, 60), dtype=torch.float32, device=dev)
        grid_t = _seed.fake_sample_grid()
        def _fill():
            _nat.check(env_t._lib.ss_create_temp_states(env_t._h, 5, grid_t.ctypes.data, None, out_t.data_ptr(), env_t._stream()), "ss_create_temp_states")
        for _ in range(3):
            _fill()
        torch.cuda.synchronize(dev)
        ts_ms = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _fill(); b.record(); torch.cuda.synchronize(dev)
            ts_ms.append(a.elapsed_time(b))
        ts_s = float(np.mean(ts_ms)) * 1e-3
        tsf = {"envs": 48, "cells": cells, "bytes": 48 * cells * 240, "avg_launch_ms": ts_s * 1e3,
               "achieved": 48 * cells * 240 / ts_s / 1e9}
        del out_t
        env_t.close()
    except Exception as exc:                                  # never let the side measurement break the bench line
        tsf = {"error": repr(exc)[:200]}

    exchange = None
    if world > 1 and args.rollout > 0:
        try:
            exchange = exchange_block(env, dev, world, rank, N, args.rollout, acts)
        except Exception as exc:
            exchange = {"error": repr(exc)[:300]}

    peak, peak_src = measured_peaks()
    kern_s = t_dev / K
    achieved = ALG_BYTES_PER_ENV_STEP * N / kern_s / 1e9
    # DRAM traffic and the compute-side numbers come from ONE committed `ncu --set full` capture of this kernel
    # (tools/final_evidence.sh -> tools/ncu_summary.py -> profiles/traffic.json, which names the capture and the git head it
    # was taken at): nothing measured under a profiler is a bench value, and a stale capture is visible as such.
    traffic, compute = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("env_step_kernel_dram_bytes_per_launch")
        compute = {"kernel": "env_step_kernel", "bound": "fp32 issue / dependent latency (SURVEY 8d)",
                   "issue_active_pct": tj.get("issue_active_pct"), "pipe_fma_pct_of_peak": tj.get("pipe_fma_pct_of_peak"),
                   "pipe_alu_pct_of_peak": tj.get("pipe_alu_pct_of_peak"),
                   "warp_instructions_per_env_substep": tj.get("warp_instructions_per_env_substep"),
                   "traffic_over_algorithmic": tj.get("traffic_over_algorithmic"),
                   "local_memory_instructions_per_launch": (tj.get("local_load_instructions") or 0) + (tj.get("local_store_instructions") or 0),
                   "registers_per_thread": tj.get("registers_per_thread"), "stall_per_issue": tj.get("stall_per_issue"),
                   "capture": tj.get("capture"), "capture_git_head": tj.get("git_head"),
                   "note": "from the committed ncu capture (not measured in this run); issue-slot utilisation is the meaningful roofline of this kernel"}
    aba_achieved = ALG_BYTES_PER_ENV_SUBSTEP * N / aba_s / 1e9

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu, k_cpu = 4096, 384       # ~10-30 s of CPU work on a 16..32-core host
            rate, dt = cpu_port_rate(wl, n_cpu, k_cpu, 4)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                   "sample": "%d envs x %d steps of the same workload, %d OpenMP threads, %.1f s" % (n_cpu, k_cpu, cores, dt)}
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "envs_per_gpu": N, "global_envs": world * N,
                       "actions": "clip(N(0,0.22),-1,1), 64-batch ring resident in HBM", "parallelism": "env-sharded x%d" % world,
                       "l2": "flushed between timed steps (256 MiB write outside the event pair)",
                       "timing": "sum of per-step CUDA-event durations, max over ranks",
                       "wall_ms_per_step_incl_flush": t_wall / K * 1e3, "contacts_per_env": contacts_main},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": N * H2D_PER_ENV,
                    "d2h_bytes_per_step": N * D2H_PER_ENV, "steps": Ke, "api": "BatchedStepperEnv.step (numpy mode) -> ss_step_host",
                    "host_us_per_step": host_us, "frac_of_device_rate": e2e_value / value},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "env_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALG_BYTES_PER_ENV_STEP, "avg_launch_ms": kern_s * 1e3,
                         "note": "FP32-latency bound path (SURVEY 8d): the HBM fraction is reported as asked, not a target"},
            "roofline_aba": {"kernel": "aba_substep_kernel", "bound": "hbm", "achieved": aba_achieved, "peak": peak, "unit": "GB/s",
                             "frac": aba_achieved / peak, "traffic": None, "avg_launch_ms": aba_s * 1e3,
                             "algorithmic_bytes_per_env_substep": ALG_BYTES_PER_ENV_SUBSTEP,
                             "env_substeps_per_s": N / aba_s,
                             "large_batch": {"envs": NL, "avg_launch_ms": abaL_s * 1e3,
                                             "achieved": ALG_BYTES_PER_ENV_SUBSTEP * NL / abaL_s / 1e9,
                                             "frac": ALG_BYTES_PER_ENV_SUBSTEP * NL / abaL_s / 1e9 / peak}},
        }
        if compute is not None:
            line["roofline_compute"] = compute
        if tsf is not None:
            if "achieved" in tsf:
                tsf.update(kernel="temp_prep_kernel + temp_fill_kernel", bound="hbm", peak=peak, unit="GB/s", frac=tsf["achieved"] / peak,
                           algorithmic_bytes_per_env_cell=240)
            line["roofline_temp_states"] = tsf
        if walking is not None:
            line["walking"] = walking
        if exchange is not None:
            line["exchange"] = exchange
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    ap.add_argument("--workload", default="w3d-flat", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-envs", type=int, default=0, help="envs per step of the CPU arm (0 = the GPU arm's --envs: same config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-walking", action="store_true", help="skip the secondary measurement on walking states")
    ap.add_argument("--rollout", type=int, default=2048,
                    help="N > 1 only: length T of the rollout whose exchange step (configs[3]) is timed after the step metric; 0 = skip")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
