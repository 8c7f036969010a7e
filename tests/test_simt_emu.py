"""CPU: the PRODUCT's CUDA sources (steppingstone_b200/csrc/ss_kernels.cu + ss_device.cuh), compiled for the host on the
SIMT emulation of tests/simt_emu/ (one fiber per CUDA thread, warp collectives as rendez-vous), against the oracle.

This is test infrastructure: it checks the kernels' arithmetic and warp-level data flow in the CPU suite (where no GPU
exists) and lets a kernel change be verified before it is timed on a B200.  It is not a CPU path of the product -- the
package cannot load this library, and the `-m gpu` tests remain the parity tests proper (same comparisons, real device).
Differences from the device: libm sqrt/sincos/exp instead of the GPU's approximations, no FMA contraction of a*b+c;
both are far below the fp32-vs-fp64 tolerances used here (the same as in tests/test_gpu_parity.py).
"""
import os

import numpy as np
import pytest

from conftest import make_oracle, random_actions
from simt_emu.binding import EmuEnv

KINDS = [0, 1]


def _maxerr(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


@pytest.mark.parametrize("kind", KINDS)
def test_emulated_kernels_reset_parity(kind):
    """env_reset_kernel: terrain generator on the env's MT19937 stream (bit-exact), planks, bookkeeping, observation."""
    n = 6
    o = make_oracle(kind, n); g = EmuEnv(kind, n)
    obs_o = o.reset(); obs_g = g.reset()
    so, sg = o.get_state(), g.get_state()
    assert np.array_equal(so["mt"], sg["mt"])
    assert np.array_equal(so["ints"], sg["ints"])
    assert _maxerr(so["terrain"], sg["terrain"]) < 1e-9
    assert _maxerr(so["plank"], sg["plank"]) < 1e-9
    assert _maxerr(so["q"], sg["q"]) < 1e-6
    assert _maxerr(obs_o, obs_g) < 1e-5
    o.close(); g.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("k_steps", [0, 12])
def test_emulated_kernels_teacher_forced_substeps(kind, k_steps):
    """env_step_kernel, per substep (dynamics + collision + rows + 20 PGS iterations + integration) from the oracle's
    state: the comparison and tolerances of tests/test_gpu_parity.py::test_teacher_forced_substeps."""
    n = 8
    prm = dict(n_substeps=1)
    o = make_oracle(kind, n, params=prm); g = EmuEnv(kind, n, params=prm)
    o.reset(); g.reset()
    rng0 = np.random.RandomState(k_steps)
    for _ in range(4 * k_steps):
        o.step(random_actions(rng0, n))
    g.enable_substep_trace(True)
    rng = np.random.RandomState(100 + k_steps)
    worst_v = worst_q = 0.0
    for s in range(6):
        a = random_actions(rng, n)
        st = o.get_state(); g.set_state(st)
        o.step(a); g.step(a)
        qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
        vscale = np.maximum(1.0, np.abs(qdo[0]).max(axis=1, keepdims=True))
        worst_v = max(worst_v, float(np.max(np.abs(qdo[0] - qdg[0]) / vscale))); worst_q = max(worst_q, _maxerr(qo[0], qg[0]))
    o.close(); g.close()
    assert worst_v < 1e-3 and worst_q < 5e-6, (worst_v, worst_q)


@pytest.mark.parametrize("kind", KINDS)
def test_emulated_kernels_chained_control_step_is_strict(kind):
    """One whole control step (4 chained substeps) of 96 envs from oracle states, the strict criterion of
    tests/parity_helpers.py::chain_check: every env whose substeps took the oracle's discrete decisions (limit rows, contact
    points, velocity clamp: ss_substep_events) is within the fp32 tolerance at every substep; every flip is accounted for."""
    from parity_helpers import chain_check
    n = 96
    o = make_oracle(kind, n); g = EmuEnv(kind, n)
    o.reset(); g.reset()
    rng = np.random.RandomState(40 + kind)
    for _ in range(25):
        o.step(random_actions(rng, n))
    g.enable_substep_trace(True)
    for t in range(2):
        st = o.get_state(); g.set_state(st)
        a = random_actions(rng, n)
        obs_o, rew_o, done_o, *_ = o.step(a); obs_g, rew_g, done_g, *_ = g.step(a)
        strict = chain_check(o, o.substep_trace(), g.substep_trace(), o.substep_events(), g.substep_events(), st["q"],
                             "emu kind %d step %d" % (kind, t))
        assert strict.mean() >= 0.75
        assert np.array_equal(done_o.astype(bool)[strict], done_g.astype(bool)[strict])
        sel = strict & ~done_g.astype(bool)
        cols = np.ones(60, bool); cols[6:9] = kind == 0
        assert np.abs(obs_o - obs_g)[sel][:, cols].max() < 5e-3
    o.close(); g.close()


_EMU_GOLDEN = ["w3d_sim_seed8", "w3d_ghost_laststone", "w3d_walk_policy", "mike_sim_seed9", "mike_ghost_sampleprob",
               "mike_ghost_sampleprob_at_done", "mike_ghost_timelimit", "mike_walk_policy"]


@pytest.mark.parametrize("name", _EMU_GOLDEN)
def test_emulated_kernels_replay_reference_goldens(name):
    """The REFERENCE-recorded tapes (tools/gen_golden.py: the unmodified env classes inside the reference's worker wrappers)
    through the product's kernels -- ss_set_state + ss_step_injected on the emulation -- at the oracle's tolerances
    (observation 2e-6 relative, reward 1e-5): the env logic of the CUDA sources is pinned to the reference directly, not
    only through the oracle.  (`-m gpu` runs all 13 tapes on the device: test_gpu_replays_reference_goldens.)"""
    from conftest import ROOT
    from parity_helpers import golden_params, replay_golden
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    b = EmuEnv(int(g["kind"]), 1, seed=int(g["seed"]), params=golden_params(g))
    assert replay_golden(g, b, terrain_tol=1e-9, pot_tol=1e-7) == len(g["rew"])
    b.close()


@pytest.mark.parametrize("case", ["w3d_latest_shipped_terrain", "mike_latest_curriculum5"])
def test_emulated_kernels_on_walking_states(case):
    """Per-substep parity on snapshots of a shipped policy walking the stones (feet loaded, self-contacts, hinges on
    limits) -- the CPU twin of tests/test_policy_behaviour.py::test_per_substep_parity_on_walking_states."""
    from steppingstone_b200 import seeding
    from test_policy_behaviour import CASES, actor_numpy, load_actor
    name, kind, terrain_mode, curriculum = CASES[case]
    n = 8
    P = load_actor(name)
    walker = make_oracle(kind, n, terrain_mode=terrain_mode)
    if curriculum is not None:
        walker.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(curriculum)))
    prm = dict(n_substeps=1)
    o = make_oracle(kind, n, params=prm, terrain_mode=terrain_mode)
    g = EmuEnv(kind, n, params=prm, terrain_mode=terrain_mode)
    o.reset(); g.reset(); g.enable_substep_trace(True)
    obs = walker.reset()
    worst_v = worst_q = 0.0
    n_contacts = 0
    for t in range(1, 251):
        a = actor_numpy(P, obs)
        if t in (40, 130, 250):
            o.set_state(walker.get_state())
            for s in range(4):
                st = o.get_state(); g.set_state(st)
                n_contacts += int((st["warm"][:, :8] >= 0).sum())
                o.step(a); g.step(a)
                qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
                vscale = np.maximum(1.0, np.abs(qdo[0]).max(axis=1, keepdims=True))
                worst_v = max(worst_v, float(np.max(np.abs(qdo[0] - qdg[0]) / vscale)))
                worst_q = max(worst_q, _maxerr(qo[0], qg[0]))
        obs = walker.step(a)[0]
    walker.close(); o.close(); g.close()
    assert n_contacts > 50
    assert worst_v < 1e-3 and worst_q < 5e-6, (worst_v, worst_q)


def test_product_loader_refuses_the_emulated_library():
    """Even the experiment hook of the product's loader (SS_B200_LIB, a differently-tuned CUDA build) will not take it."""
    import subprocess
    import sys
    from simt_emu import build_emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SS_B200_LIB=build_emu.build())
    code = "import sys; sys.path.insert(0, %r); from steppingstone_b200 import native; native.load()" % root
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode != 0 and "SIMT emulation" in res.stderr


def test_emulation_is_not_reachable_from_the_product():
    """The emulated library lives under tests/ and has its own binding: the package never names it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "steppingstone_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libss_emu" not in src and "simt_emu.binding" not in src and "import simt_emu" not in src, f


@pytest.mark.parametrize("kind", KINDS)
def test_results_do_not_depend_on_the_lane_schedule(kind):
    """Threads of the kernels may only exchange data across warp collectives / block barriers.  The emulation resumes a
    block's fibers in ascending, descending or a shuffled order per scheduling round: reset + control steps in contact
    must be BIT-identical under all three (a kernel leaning on lock-step execution between two collectives would not be)."""
    from simt_emu import binding
    lib = binding.load()
    n = 5
    out = []
    warm = make_oracle(kind, n)
    warm.reset()
    rng = np.random.RandomState(3)
    for _ in range(14):
        warm.step(random_actions(rng, n))                       # falling / first contacts
    st0 = warm.get_state(); warm.close()
    acts = [random_actions(rng, n) for _ in range(3)]
    try:
        for order in (0, 1, 2):
            lib.simt_emu_set_order(order)
            g = EmuEnv(kind, n)
            obs0 = g.reset()
            g.set_state(st0)
            res = [g.step(a) for a in acts]
            out.append((obs0, [r[0] for r in res], [r[1] for r in res], g.get_state()))
            g.close()
    finally:
        lib.simt_emu_set_order(0)
    for other in out[1:]:
        assert np.array_equal(out[0][0], other[0])
        for a, b in zip(out[0][1], other[1]):
            assert np.array_equal(a, b)
        for a, b in zip(out[0][2], other[2]):
            assert np.array_equal(a, b)
        for k in ("q", "qd", "warm", "terrain", "scalars", "ints", "mt"):
            assert np.array_equal(out[0][3][k], other[3][k]), k


@pytest.mark.parametrize("kind", KINDS)
def test_emulated_aba_substep_kernel(kind):
    """aba_substep_kernel (the metric's "ABA kernel"): tolerances of tests/test_gpu_parity.py::test_aba_substep_parity."""
    n = 24
    o = make_oracle(kind, 1); g = EmuEnv(kind, 2)
    rng = np.random.RandomState(2)
    q = np.zeros((n, 28)); qd = rng.randn(n, 27)
    q[:, :3] = rng.randn(n, 3)
    quat = rng.randn(n, 4); q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, 7:] = rng.uniform(-0.5, 0.5, (n, 21))
    tau = rng.randn(n, 21) * 20
    qo, qdo = o.aba_substep(q, qd, tau)
    qg, qdg = g.aba_substep(q, qd, tau)
    assert np.max(np.abs(qdo - qdg) / np.maximum(1.0, np.abs(qdo))) < 5e-4
    assert _maxerr(qo, qg) < 1e-5
    o.close(); g.close()


def test_emulated_gae_kernel():
    T, N = 19, 70
    rng = np.random.RandomState(0)
    o = make_oracle(0, 1); g = EmuEnv(0, 1)
    r = rng.randn(T, N).astype(np.float32); v = rng.randn(T + 1, N).astype(np.float32)
    m = (rng.rand(T + 1, N) > 0.1).astype(np.float32); bm = (rng.rand(T + 1, N) > 0.05).astype(np.float32)
    assert np.max(np.abs(g.gae(r, v, m, bm, 0.99, 0.95)[:T] - o.gae(r, v, m, bm, 0.99, 0.95)[:T])) < 1e-4
    o.close(); g.close()


@pytest.mark.parametrize("kind,variant", [(0, 1), (1, 2)])
def test_emulated_temp_state_kernels(kind, variant):
    """temp_prep_kernel + temp_fill_kernel (create_temp_states_1 / _2) incl. the zero rows and the side effects."""
    from steppingstone_b200 import seeding
    n = 6
    o = make_oracle(kind, n); g = EmuEnv(kind, n)
    if kind == 1:
        cdf = seeding.choice_cdf(seeding.curriculum_table(5)); o.set_sample_cdf(cdf); g.set_sample_cdf(cdf)
    o.reset(); g.reset()
    st = o.get_state()
    # as tests/test_gpu_parity.py::_force_advance: a foot contact on the target with the dwell counter one short
    st["ints"][:, 1] = 29; st["scalars"][:, 14:16] = 1.0
    tgt = st["terrain"][np.arange(n), st["ints"][:, 0]]
    st["q"][:, 0:2] = tgt[:, 0:2]; st["q"][:, 2] = tgt[:, 2] + 0.9
    st["ints"][::3, 1] = 3                                         # every third env does NOT advance -> zero rows
    o.set_state(st); g.set_state(st)
    a = np.zeros((n, 21), np.float32)
    o.env_post(a); g.env_post(a)
    ts_o = o.create_temp_states(variant, seeding.fake_sample_grid(), seeding.temp_state_restore(kind))
    ts_g = g.create_temp_states(variant, seeding.fake_sample_grid(), seeding.temp_state_restore(kind))
    assert ts_o.shape == ts_g.shape
    assert not ts_g[::3].any() and np.abs(ts_o[1]).max() > 0
    cols = np.ones(60, bool)
    if kind == 1:
        cols[6:9] = False
    assert np.max(np.abs(ts_o - ts_g)[:, :, cols]) < 2e-4
    assert _maxerr(o.get_state()["terrain"], g.get_state()["terrain"]) < 1e-9
    o.close(); g.close()


def test_emulated_resample_kernel():
    """resample_kernel (update_sample_prob's immediate re-draw for the envs that just advanced a stone) + the later draws
    from the new table: CPU twin of tests/test_gpu_parity.py::test_update_sample_prob_resamples_flagged_envs -- the test
    that found the stale-scratch bug of this kernel on the device."""
    from steppingstone_b200 import seeding
    n = 10
    o = make_oracle(1, n); g = EmuEnv(1, n)
    cdf5 = seeding.choice_cdf(seeding.curriculum_table(5))
    o.set_sample_cdf(cdf5); g.set_sample_cdf(cdf5)
    o.reset(); g.reset()

    def force_advance(st):
        st["ints"][:, 1] = 29; st["scalars"][:, 14:16] = 1.0
        tgt = st["terrain"][np.arange(n), st["ints"][:, 0]]
        st["q"][:, 0:2] = tgt[:, 0:2]; st["q"][:, 2] = tgt[:, 2] + 0.9

    st = o.get_state(); force_advance(st)
    st["ints"][::2, 1] = 3                       # every other env does not advance: no immediate re-sample there
    o.set_state(st); g.set_state(st)
    a = np.zeros((n, 21), np.float32)
    o.env_post(a); g.env_post(a)
    rng = np.random.RandomState(12)
    table = (rng.rand(11, 11, 11, 11, 11) * (rng.rand(11, 11, 11, 11, 11) < 0.02)).astype(np.float32).astype(np.float64)
    table[5, 5, 0, 5, 5] += 1e-3
    before = o.get_state()["terrain"].copy()
    cdf = seeding.choice_cdf(table)
    o.set_sample_cdf(cdf, apply_now=True); g.set_sample_cdf(cdf, apply_now=True)
    so, sg = o.get_state(), g.get_state()
    changed = np.abs(so["terrain"] - before).max(axis=(1, 2)) > 0
    assert changed[1::2].all() and not changed[::2].any()
    assert np.array_equal(so["mt"], sg["mt"]) and np.array_equal(so["ints"], sg["ints"])
    assert _maxerr(so["terrain"], sg["terrain"]) < 1e-9 and _maxerr(so["plank"], sg["plank"]) < 1e-9
    st = o.get_state(); force_advance(st); o.set_state(st); g.set_state(st)
    o.env_post(a); g.env_post(a)
    so, sg = o.get_state(), g.get_state()
    assert np.array_equal(so["mt"], sg["mt"]) and _maxerr(so["terrain"], sg["terrain"]) < 1e-9
    o.close(); g.close()


@pytest.mark.parametrize("kind", KINDS)
def test_emulated_whole_steps_with_auto_reset(kind):
    """Whole control steps (4 chained substeps, observation, reward, termination, Monitor bookkeeping) and the auto-reset
    inside the step kernel (terrain generator on the env's MT19937 stream), each step teacher-forced from the oracle's
    state so that the chaotic contact dynamics cannot drift: episode ends agree, the post-reset observation and the RNG
    streams of the envs that ended are identical."""
    n = 5
    o = make_oracle(kind, n); g = EmuEnv(kind, n)
    o.reset(); g.reset()
    rng = np.random.RandomState(5)
    n_done = n_mismatch = 0
    for t in range(30):
        a = random_actions(rng, n, sd=1.0)                      # saturated torques: the robots fall within ~25 steps
        g.set_state(o.get_state())
        obs_o, rew_o, done_o, to_o, er_o, el_o = o.step(a)
        obs_g, rew_g, done_g, to_g, er_g, el_g = g.step(a)
        done_o = done_o.astype(bool)
        n_mismatch += int((done_o != done_g).sum())
        both = done_o & done_g
        if both.any():
            n_done += int(both.sum())
            so, sg = o.get_state(), g.get_state()
            assert np.array_equal(so["mt"][both], sg["mt"][both]), "RNG streams differ after the auto-reset"
            assert _maxerr(so["terrain"][both], sg["terrain"][both]) < 1e-9
            assert _maxerr(obs_o[both], obs_g[both]) < 1e-5, "post-reset observation"
            assert np.array_equal(el_o[both], el_g[both])
            assert _maxerr(er_o[both], er_g[both]) < 1e-2 * max(1.0, float(np.abs(er_o[both]).max()))
        calm = ~done_o & ~done_g
        assert np.abs(rew_o - rew_g)[calm].max(initial=0.0) < 0.5      # 4 chained substeps in fp32, saturated torques
    assert n_done >= 3 and n_mismatch <= 1, (n_done, n_mismatch)
    o.close(); g.close()


def test_shipped_policy_walks_on_the_emulated_kernels():
    """The reference's flat-stage policy (playground/models/..._base.pt) driving the PRODUCT's kernels closed-loop for 450
    control steps: ~14 stones per env (a stone needs 30 dwell frames) -- tests/test_policy_behaviour.py on the device code
    itself, without a device."""
    from test_policy_behaviour import actor_numpy, load_actor
    n = 3
    P = load_actor("w3d_base.npz")
    g = EmuEnv(0, n, terrain_mode=2)
    obs = g.reset()
    total = np.zeros(n)
    for t in range(450):
        obs, rew, done, _, _, _ = g.step(actor_numpy(P, obs))
        assert not done.any(), "fell at step %d" % t
        total += rew
    stones = g.get_state()["ints"][:, 0]
    g.close()
    assert (stones >= 12).all() and (total > 1500).all(), (stones, total)


@pytest.mark.parametrize("kind", KINDS)
def test_dense_factorisation_variant_stays_correct(kind):
    """The default build factorises H = U D U^T leaves-first with the structural zeros of the kinematic tree skipped at compile
    time (0.7207 -> 0.6942 ms per 8192-env step on a B200, profiles/r02b_variants.txt); -DSS_SPARSE_LDL=0 keeps the dense
    row-ordered LDL^T for A/B runs.  Both get the same per-substep comparison against the oracle, in contact; the sparse one
    needs fewer warp collectives."""
    n = 6
    prm = dict(n_substeps=1)
    o = make_oracle(kind, n, params=prm)
    g = EmuEnv(kind, n, params=prm); d = EmuEnv(kind, n, params=prm, variant="dense")
    o.reset(); g.reset(); d.reset()
    rng0 = np.random.RandomState(7)
    for _ in range(40):
        o.step(random_actions(rng0, n))
    g.enable_substep_trace(True); d.enable_substep_trace(True)
    rng = np.random.RandomState(107)
    worst_v = worst_q = 0.0
    count = {}
    for s in range(6):
        a = random_actions(rng, n)
        st = o.get_state(); g.set_state(st); d.set_state(st)
        o.step(a)
        for name, env in (("sparse", g), ("dense", d)):
            c0 = env.lib.simt_emu_collectives(); env.step(a); count[name] = count.get(name, 0) + env.lib.simt_emu_collectives() - c0
        qo, qdo = o.substep_trace()
        vscale = np.maximum(1.0, np.abs(qdo[0]).max(axis=1, keepdims=True))
        for env in (g, d):
            qg, qdg = env.substep_trace()
            worst_v = max(worst_v, float(np.max(np.abs(qdo[0] - qdg[0]) / vscale))); worst_q = max(worst_q, _maxerr(qo[0], qg[0]))
    qa = np.random.RandomState(2).uniform(-0.5, 0.5, (8, 28)); qa[:, 3:7] /= np.linalg.norm(qa[:, 3:7], axis=1, keepdims=True)
    qd = np.random.RandomState(3).randn(8, 27); tau = np.random.RandomState(4).randn(8, 21) * 20
    ref = make_oracle(kind, 1)
    qo2, qdo2 = ref.aba_substep(qa, qd, tau); qg2, qdg2 = d.aba_substep(qa, qd, tau)
    o.close(); g.close(); d.close(); ref.close()
    assert worst_v < 1e-3 and worst_q < 5e-6, (worst_v, worst_q)
    assert np.max(np.abs(qdo2 - qdg2) / np.maximum(1.0, np.abs(qdo2))) < 5e-4 and _maxerr(qo2, qg2) < 1e-5
    assert count["sparse"] < 0.92 * count["dense"], count          # ~145 of ~1300 collectives per env-substep gone


def test_production_cta_shape_on_the_emulation():
    """The shipped launch configuration -- 20 warps per CTA, ghost warps in the last CTA that only keep the per-substep
    block barrier company, envs handed to warps in contact-count order (sort kernels + perm) -- with 85 envs = 4 full CTAs
    + 5 envs and 15 ghost warps: identical results to the one-warp-per-CTA emulation build (which warp steps which env
    never changes a result), under two fiber schedules."""
    n = 85
    o = make_oracle(0, n)
    o.reset()
    rng = np.random.RandomState(9)
    for _ in range(12):
        o.step(random_actions(rng, n))
    st = o.get_state(); o.close()
    acts = [random_actions(rng, n) for _ in range(2)]
    res = {}
    for variant, order in (("", 0), ("wpb20", 0), ("wpb20", 1)):
        g = EmuEnv(0, n, variant=variant)
        g.lib.simt_emu_set_order(order)
        try:
            g.reset(); g.set_state(st)
            out = [g.step(a) for a in acts]
            res[(variant, order)] = (out, g.get_state())
        finally:
            g.lib.simt_emu_set_order(0)
            g.close()
    ref_out, ref_st = res[("", 0)]
    for key in (("wpb20", 0), ("wpb20", 1)):
        out, stt = res[key]
        for a, b in zip(ref_out, out):
            for x, y in zip(a, b):
                assert np.array_equal(x, y), key
        for k in ("q", "qd", "warm", "scalars", "ints", "mt", "terrain"):
            assert np.array_equal(ref_st[k], stt[k]), (key, k)


def test_plain_c_host_runs_against_the_emulated_library(tmp_path):
    """examples/c_host.c (the host-without-Python example of INTEGRATION.md) linked against the emulated library instead of
    libss_b200.so: the C-ABI of include/ss_env.h end to end from plain C -- ss_create, ss_reset_host, ss_step_host with
    host buffers, ss_destroy -- the CPU twin of tests/test_abi.py::test_c_host_runs_on_gpu."""
    import subprocess
    import sys
    from simt_emu import build_emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = build_emu.build()
    exe = str(tmp_path / "c_host_emu")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "examples", "c_host.c"), lib, "-Wl,-rpath," + os.path.dirname(lib), "-lm"])
    tables = str(tmp_path / "tables")
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "dump_tables.py"), tables, "walker3d", "8", "8"],
                          stdout=subprocess.DEVNULL)
    res = subprocess.run([exe, tables, "8", "30"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    assert "env-steps/s" in res.stdout and "episodes finished" in res.stdout
