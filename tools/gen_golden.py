"""Generate tests/golden/*.npz by running the REFERENCE's own env code (build container only).

    python tools/gen_golden.py            # reads /root/reference, writes tests/golden/

What is pinned.  The reference's stepping-stone envs are Python on top of PyBullet.  PyBullet (third party,
unpinned, requirements.txt:3) is not in the reference tree nor in this image, so its arithmetic cannot be
run; everything ABOVE it can.  This script imports mocca_envs.walker3d_envs.Walker3DStepperEnv and
mocca_envs.mike.MikeStepperEnv *unmodified* with tools/fakes/{pybullet,gym} on the path, wraps them exactly as
a ShmemVecEnv worker does (gym TimeLimit 1000 -> TimeLimitMask -> Monitor, common/envs_utils.py:25-40; auto
reset :645-649) and drives them with "physics tapes": per control step the post-physics state (q, qd, foot
contact flags) that stepSimulation would have produced is installed through the fake's hooks (foot link
positions come from the oracle's forward kinematics).  Recorded: what the reference then computes --
observation, reward, done, bad_transition, Monitor episode stats, next_step_index, dwell counter, terrain_info.
tests/test_oracle_golden.py replays the same tapes through the oracle's env logic (sso_step_injected) and must
reproduce these outputs; the GPU path is then checked against the oracle (tests/test_gpu_parity.py).

Tapes: (A) oracle-simulated random-action rollouts (falls, resets, RNG continuity); (B) a kinematic "ghost
walker" that dwells on each stone (stone advances, step bonus, plank ring buffer, curriculum sampler, time-limit
truncation); (C) the same with stop_frames=5 and no time limit, to reach the last stone (target bonus, index clamp);
(D) ghost tapes that call create_temp_states_1/_2/create_temp_states after every stone advance; (E) a Mike ghost tape that
calls update_sample_prob with a fresh sparse 11^5 table after stone advances (immediate re-sample + later draws);
(G) the same with a 23-step time limit, so that episodes END on stone-advance steps: reset() leaves update_terrain
set and the following update_sample_prob acts on the freshly reset env; (F) the reference's shipped policies walking the stones in the oracle's simulation (closed loop), replayed through the
reference env: stone advances by physical foot contact, step bonuses and progress rewards on realistic gaits.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tools", "fakes"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import pybullet as fake  # noqa: E402  (tools/fakes/pybullet.py)
import gym  # noqa: E402  (tools/fakes/gym)
from gym.wrappers import TimeLimit  # noqa: E402

from common.envs_utils import Monitor, TimeLimitMask  # noqa: E402  (reference)
from mocca_envs.mike import MikeStepperEnv  # noqa: E402  (reference)
from mocca_envs.walker3d_envs import Walker3DStepperEnv  # noqa: E402  (reference)

from oracle.binding import Oracle  # noqa: E402
from steppingstone_b200 import model, seeding  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_oracle(kind, seed, params=None):
    M = model.make_model(kind)
    P = model.make_params(kind, params)
    return Oracle(M, P, 1, seeding.env_streams(seed, 0, 1), seeding.aux_streams(kind, seed, 0, 1), threads=1)


class RefWorker:
    """One reference env wrapped and driven like common/envs_utils._subproc_worker."""

    def __init__(self, kind, seed, fk_oracle, time_limit=1000, env_params=None, curriculum=None):
        fake.reset_world()
        self.kind = kind
        np.random.seed(seed)                                   # seed_all in the parent (common/sacred_utils.py:64-71)
        env = (Walker3DStepperEnv if kind == 0 else MikeStepperEnv)()
        env.seed(seed)                                         # make_env_fns: env.seed(seed + rank)
        # The robot keeps the RandomState made by seed(None) in EnvBase.__init__ (entropy).  Pin it to the
        # documented aux stream so that the Mike mirror coin is reproducible (steppingstone_b200/seeding.py).
        if kind == 1:
            env.robot.np_random = seeding.rng_from_words(seeding.aux_streams(1, seed, 0, 1)[0])
        if env_params:
            env.set_env_params(env_params)                     # env_base.py:73-76
        if curriculum is not None:
            env.update_curriculum(curriculum)
        self.env = env
        w = env
        if time_limit:
            w = TimeLimitMask(TimeLimit(env, max_episode_steps=time_limit))
        self.wrapped = Monitor(w, None, allow_early_resets=True)
        self.robot_id = env.robot.object_id[0]
        self.foot_links = [f.bodyPartIndex for f in env.robot.feet]
        self.fk = fk_oracle
        self.contact = np.zeros(2)
        self.pending = None
        fake.WORLD["link_pose_hook"] = self._link_pose
        fake.WORLD["contact_hook"] = self._contacts
        fake.WORLD["step_hook"] = self._install

    def _q(self):
        b = fake.WORLD["bodies"][self.robot_id]
        hinges = [j["q"] for j in b.joints if not j["name"].startswith("jointfix")]
        return np.array(list(b.pos) + list(b.orn) + hinges)

    def _link_pose(self, body, link):
        if body == self.robot_id and link in self.foot_links:
            return tuple(self.fk.feet_fk(self._q())[self.foot_links.index(link)])
        return None

    def _contacts(self, body, link):
        if body == self.robot_id and link in self.foot_links and self.contact[self.foot_links.index(link)] != 0:
            return ((0, body, 4242, link, -1),)
        return ()

    def _install(self):
        q, qd, c = self.pending
        b = fake.WORLD["bodies"][self.robot_id]
        b.pos, b.orn = list(q[:3]), list(q[3:7])
        b.ang, b.lin = list(qd[:3]), list(qd[3:6])
        k = 0
        for j in b.joints:
            if not j["name"].startswith("jointfix"):
                j["q"], j["qd"] = float(q[7 + k]), float(qd[6 + k]); k += 1
        self.contact = np.array(c, dtype=float)

    def reset(self):
        return np.asarray(self.wrapped.reset(), dtype=np.float32)

    def step(self, a, q, qd, c):
        self.pending = (q, qd, c)
        ob, r, d, info = self.wrapped.step(a)
        internals = (self.env.next_step_index, self.env.target_reached_count, float(self.env.linear_potential))
        if d:
            ob = self.wrapped.reset()                          # envs_utils.py:646-649
        return np.asarray(ob, dtype=np.float32), float(r), bool(d), info, internals


def record(name, kind, seed, actions, tape_fn, T, time_limit=1000, env_params=None, curriculum=None, oracle_params=None,
           temp_states=False, sample_prob_every=0, sample_prob_at_done=False, sample_prob_unflagged_every=0):
    """tape_fn(t, worker, sim) -> (q, qd, contact) for step t.  temp_states: after every step that advanced a stone
    call the reference's create_temp_states_1 / _2 / create_temp_states (value-guided sampling probes) and record them."""
    fk = make_oracle(kind, seed)
    w = RefWorker(kind, seed, fk, time_limit, env_params, curriculum)
    rec = dict(kind=kind, seed=seed, time_limit=time_limit or 0, curriculum=-1 if curriculum is None else curriculum,
               stop_frames=(env_params or {}).get("stop_frames", 30))
    obs0 = w.reset()
    terr = [w.env.terrain_info.copy()]; terr_at = [-1]
    Q, QD, C, OBS, REW, DONE, BAD, EPR, EPL, NXT, CNT, POT = ([] for _ in range(12))
    ts_step, ts_variant, ts_data = [], [], {}
    sp_rng = np.random.RandomState(4242 + seed); sp_step, sp_tables = [], []
    for t in range(T):
        q, qd, c = tape_fn(t, w)
        ob, r, d, info, (nxt, cnt, pot) = w.step(actions[t], q, qd, c)
        Q.append(q); QD.append(qd); C.append(c); OBS.append(ob); REW.append(r); DONE.append(d)
        BAD.append(bool(info.get("bad_transition", False)))
        EPR.append(info["episode"]["r"] if "episode" in info else 0.0)
        EPL.append(info["episode"]["l"] if "episode" in info else 0)
        NXT.append(nxt); CNT.append(cnt); POT.append(pot)
        unflagged = bool(sample_prob_unflagged_every) and not w.env.update_terrain and t % sample_prob_unflagged_every == 3
        if unflagged or (sample_prob_every and (not d or sample_prob_at_done) and w.env.update_terrain and (d or (nxt % sample_prob_every) != 0)):
            # update_sample_prob (mike.py:827-830 through envs_utils.py:568-571) right after a step that advanced a stone
            # (the only moment the reference env takes the table): new sampling table + immediate re-sample of the
            # next-next stone.  Skipped on every third stone so that draws from a table set earlier are covered too.
            # sample_prob_at_done: also after an episode that ENDED on an advance step -- reset() does not clear
            # update_terrain, so the freshly reset env takes the table and re-samples its stone 2 (tape G).
            # sample_prob_unflagged_every: the table is ALSO sent at steps that advanced nothing -- the reference env
            # ignores it then (`if self.update_terrain:`, mike.py:828) and keeps drawing from the table it has (tape G).
            table = sp_rng.rand(11, 11, 11, 11, 11) * (sp_rng.rand(11, 11, 11, 11, 11) < 0.02)
            table[5, 5, 0, 5, 5] += 1e-3
            table = table.astype(np.float32).astype(np.float64)          # stored as float32 in the fixture: keep it exact
            w.env.update_sample_prob(table.copy())
            sp_step.append(t); sp_tables.append(table)
        if temp_states and not d and w.env.update_terrain:
            k = len(ts_step)
            variant = 5 if k == 2 else (2 if k % 4 == 1 and k < 10 else 1)
            fn = {1: w.env.create_temp_states_1, 2: w.env.create_temp_states_2, 5: w.env.create_temp_states}[variant]
            ts = np.asarray(fn())
            ts_step.append(t); ts_variant.append(variant)
            if variant == 5:                                   # 161051 x 60: keep a strided sample + column sums
                ts_data["ts_%d" % k] = ts[::1009].astype(np.float32)
                ts_data["ts_sum_%d" % k] = ts.sum(axis=0)
            else:
                ts_data["ts_%d" % k] = ts.astype(np.float32)
        if d or t == T - 1 or (t % 97 == 0):
            terr.append(w.env.terrain_info.copy()); terr_at.append(t)
    rec.update(actions=np.asarray(actions[:T], np.float32), q=np.array(Q), qd=np.array(QD), contact=np.array(C),
               obs0=obs0, obs=np.array(OBS, np.float32), rew=np.array(REW), done=np.array(DONE), bad=np.array(BAD),
               ep_r=np.array(EPR), ep_l=np.array(EPL, np.int32), next_step=np.array(NXT, np.int32),
               reached_count=np.array(CNT, np.int32), potential=np.array(POT),
               terrain=np.array(terr), terrain_at=np.array(terr_at, np.int32))
    if sample_prob_every:
        rec.update(sp_step=np.array(sp_step, np.int32), sp_tables=np.array(sp_tables, np.float32))
    if temp_states:
        e = w.env
        rec.update(ts_step=np.array(ts_step, np.int32), ts_variant=np.array(ts_variant, np.int32),
                   ts_grid=np.array([e.fake_yaw_samples, e.fake_pitch_samples, e.fake_r_samples, e.fake_x_tilt_samples,
                                     e.fake_y_tilt_samples]), **ts_data)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-28s T=%d episodes=%d max next_step=%d reward[min,max]=[%.2f, %.2f] sample_prob calls %d (%d right after a done step)" % (
        name, T, int(np.sum(DONE)), int(np.max(NXT)), np.min(REW), np.max(REW), len(sp_step), sum(bool(DONE[t]) for t in sp_step)))
    if sample_prob_unflagged_every:
        print("   ... %d of them at steps that advanced no stone (the env must ignore the table)" % sum(1 for t in sp_step if t % sample_prob_unflagged_every == 3))


def simulated_tape(kind, seed, actions, T, curriculum=None):
    """(A) post-physics states from the oracle's own simulation (normal mode)."""
    sim = make_oracle(kind, seed)
    if curriculum is not None:
        sim.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(curriculum)))
    sim.reset()

    def fn(t, w):
        sim.step(actions[t][None])
        q, qd = sim.substep_trace()
        st = sim.get_state()
        return q[3, 0].copy(), qd[3, 0].copy(), st["scalars"][0, 14:16].copy()
    return fn


def policy_actions(kind, seed, T, fixture, curriculum=None):
    """(F) action tape of a shipped reference policy (tests/golden/policies/*.npz, tools/read_legacy_checkpoint.py)
    walking the stones closed-loop in the oracle: common/controller.py SoftsignActor.forward, deterministic."""
    z = np.load(os.path.join(OUT, "policies", fixture))
    sim = make_oracle(kind, seed)
    if curriculum is not None:
        sim.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(curriculum)))
    obs = sim.reset()
    softsign = lambda x: x / (1.0 + np.abs(x))
    acts = []
    for t in range(T):
        x = softsign(obs @ z["fc1_w"].T + z["fc1_b"]); x = softsign(x @ z["fc2_w"].T + z["fc2_b"])
        x = softsign(x @ z["fc3_w"].T + z["fc3_b"])
        x = np.maximum(x @ z["fc4_w"].T + z["fc4_b"], 0.0); x = np.maximum(x @ z["fc5_w"].T + z["fc5_b"], 0.0)
        a = np.tanh(x @ z["out_w"].T + z["out_b"]).astype(np.float32)
        acts.append(a[0])
        obs = sim.step(a)[0]
    sim.close()
    return np.array(acts, np.float32)


def ghost_tape(kind, seed, rng, dwell):
    """(B/C) kinematic walker: stands with one foot near the current target for `dwell` frames, then hops on."""
    fk = make_oracle(kind, seed)
    M = model.load_tables(kind)
    lo, hi = M["q_lo"], M["q_hi"]
    state = dict(phase=0, foot=0, pose=None)

    def fn(t, w):
        env = w.env
        tgt = env.terrain_info[env.next_step_index, 0:3]
        if state["pose"] is None or state["phase"] % dwell == 0:
            foot = int(rng.randint(2))
            yaw = env.terrain_info[env.next_step_index, 3] + rng.uniform(-0.3, 0.3)
            roll, pitch = rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.5)
            quat = np.array(fake.getQuaternionFromEuler((roll, pitch, yaw)))
            qj = np.clip(M["init_q"] + rng.uniform(-0.2, 0.2, 21), lo, hi)
            if rng.rand() < 0.3:
                k = rng.randint(3 if kind == 1 else 0, 21); qj[k] = hi[k] if rng.rand() < 0.5 else lo[k]
            if kind == 1:
                qj[:3] = rng.uniform(lo[:3], hi[:3])
            q0 = np.concatenate(([0, 0, 0], quat, qj))
            off = fk.feet_fk(q0)[foot]
            miss = rng.uniform(-0.12, 0.12, 3) * [1, 1, 0.2]
            base = tgt + miss - off
            state.update(foot=foot, pose=np.concatenate((base, quat, qj)))
        state["phase"] += 1
        q = state["pose"].copy()
        q[:3] += rng.uniform(-0.002, 0.002, 3)
        qd = rng.randn(27) * 0.5
        c = np.zeros(2); c[state["foot"]] = 1.0 if rng.rand() < 0.97 else 0.0
        c[1 - state["foot"]] = float(rng.rand() < 0.3)
        return q, qd, c
    return fn


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only in (None, "mike_ghost_sampleprob_at_done"):
        T = 200
        rng = np.random.RandomState(67)
        acts = np.clip(rng.randn(T, 21) * 0.3, -1.2, 1.2).astype(np.float32)
        record("mike_ghost_sampleprob_at_done", 1, 15, acts, ghost_tape(1, 15, rng, 6), T, time_limit=23,
               env_params={"stop_frames": 5}, curriculum=5, sample_prob_every=3, sample_prob_at_done=True,
               sample_prob_unflagged_every=7)
        if only:
            return
    for kind, nm in ((0, "w3d"), (1, "mike")):
        for seed in (8, 9):
            T = 220
            rng = np.random.RandomState(1000 + seed)
            acts = np.clip(rng.randn(T, 21) * 0.5, -1.4, 1.4).astype(np.float32)     # includes |a| > 1 (clip + q5)
            record("%s_sim_seed%d" % (nm, seed), kind, seed, acts, simulated_tape(kind, seed, acts, T), T,
                   curriculum=(5 if kind == 1 else None))
        T = 1150
        rng = np.random.RandomState(77 + kind)
        acts = np.clip(rng.randn(T, 21) * 0.3, -1.2, 1.2).astype(np.float32)
        record("%s_ghost_timelimit" % nm, kind, 11, acts, ghost_tape(kind, 11, rng, 31), T,
               curriculum=(5 if kind == 1 else None))
        T = 99 * 6 + 40 if kind == 0 else 39 * 6 + 60
        rng = np.random.RandomState(99 + kind)
        acts = np.clip(rng.randn(T, 21) * 0.3, -1.2, 1.2).astype(np.float32)
        record("%s_ghost_laststone" % nm, kind, 12, acts, ghost_tape(kind, 12, rng, 6), T, time_limit=None,
               env_params={"stop_frames": 5}, curriculum=(3 if kind == 1 else None))
        if kind == 1:
            T = 160
            rng = np.random.RandomState(66)
            acts = np.clip(rng.randn(T, 21) * 0.3, -1.2, 1.2).astype(np.float32)
            record("mike_ghost_sampleprob", kind, 14, acts, ghost_tape(kind, 14, rng, 6), T, time_limit=None,
                   env_params={"stop_frames": 5}, curriculum=5, sample_prob_every=3)
        # (F) the shipped policy walking: stone advances by physical foot contact, step bonuses, progress rewards
        T = 420
        cur = 5 if kind == 1 else None
        acts = policy_actions(kind, 8, T, "w3d_latest.npz" if kind == 0 else "mike_latest.npz", cur)
        record("%s_walk_policy" % nm, kind, 8, acts, simulated_tape(kind, 8, acts, T, cur), T, curriculum=cur)
        T = 150
        rng = np.random.RandomState(55 + kind)
        acts = np.clip(rng.randn(T, 21) * 0.3, -1.2, 1.2).astype(np.float32)
        record("%s_ghost_tempstates" % nm, kind, 13, acts, ghost_tape(kind, 13, rng, 6), T, time_limit=None,
               env_params={"stop_frames": 5}, curriculum=(5 if kind == 1 else None), temp_states=True)


if __name__ == "__main__":
    main()
