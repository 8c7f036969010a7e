"""Tier-2 parity fixtures: record traces from the REAL reference (needs pybullet + gym < 0.22; neither is in the build
image, so this script has never run there -- it is the recipe SURVEY 8(c) asks for, to be run wherever PyBullet exists).

    python tools/record_pybullet_trace.py            # reads $SS_REFERENCE (default /root/reference), writes tests/golden/

For Walker3DStepperEnv-v0 and MikeStepperEnv-v0, seeds 8..15 (train.py:41 seed 8 + rank), 128 control steps of a fixed
action tape (clip(N(0, 0.22), -1, 1), numpy seed = env seed) it records, per control step: the generalised state after
the step (base position, base quaternion xyzw, 21 hinge angles; base angular + linear velocity, 21 hinge rates), the
observation, reward, done, the per-foot contact flags, next_step_index and terrain_info; once per robot: mass, local
inertia diagonal and inertial frame of every link (getDynamicsInfo) and the joint limits (getJointInfo).
tests/test_pybullet_trace.py (skipped while no pybullet_*.npz exists) teacher-forces the oracle and the GPU path from
each recorded state and compares the next one: that pins the rigid-body arithmetic, which is "parity unpinned" today.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(ROOT, "tests", "golden")


def generalised_state(env):
    p = env._p
    rid = env.robot.object_id[0]
    pos, orn = p.getBasePositionAndOrientation(rid)
    lin, ang = p.getBaseVelocity(rid)
    js = [j.get_state() for j in env.robot.ordered_joints]
    q = np.array(list(pos) + list(orn) + [s[0] for s in js])
    qd = np.array(list(ang) + list(lin) + [s[1] for s in js])
    return q, qd


def link_table(env):
    p = env._p
    rid = env.robot.object_id[0]
    rows = []
    for link in range(-1, p.getNumJoints(rid)):
        d = p.getDynamicsInfo(rid, link)
        name = p.getBodyInfo(rid)[0].decode() if link < 0 else p.getJointInfo(rid, link)[12].decode()
        rows.append((name, d[0], d[2], d[3], d[4], d[1]))          # mass, inertia diag, inertial pos, inertial orn, friction
    limits = [(j.joint_name, j.lowerLimit, j.upperLimit) for j in env.robot.ordered_joints]
    return rows, limits


def record(env_id, seed, T=128):
    import gym
    env = gym.make(env_id)
    env.seed(seed)
    base = env.unwrapped
    rng = np.random.RandomState(seed)
    acts = np.clip(rng.randn(T, 21) * 0.22, -1, 1).astype(np.float32)
    obs0 = env.reset()
    q0, qd0 = generalised_state(base)
    rec = dict(env_id=env_id, seed=seed, actions=acts, obs0=np.asarray(obs0, np.float32), q0=q0, qd0=qd0,
               terrain0=base.terrain_info.copy())
    Q, QD, OBS, REW, DONE, CON, NXT = [], [], [], [], [], [], []
    for t in range(T):
        ob, r, d, info = env.step(acts[t])
        q, qd = generalised_state(base)
        Q.append(q); QD.append(qd); OBS.append(np.asarray(ob, np.float32)); REW.append(float(r)); DONE.append(bool(d))
        CON.append(np.array(base.robot.feet_contact, dtype=np.float64)); NXT.append(int(base.next_step_index))
        if d:
            break
    rows, limits = link_table(base)
    rec.update(q=np.array(Q), qd=np.array(QD), obs=np.array(OBS), rew=np.array(REW), done=np.array(DONE),
               feet_contact=np.array(CON), next_step=np.array(NXT, np.int32), terrain=base.terrain_info.copy(),
               link_names=np.array([r[0] for r in rows]), link_mass=np.array([r[1] for r in rows]),
               link_inertia=np.array([r[2] for r in rows]), link_inertial_pos=np.array([r[3] for r in rows]),
               joint_names=np.array([l[0] for l in limits]), joint_limits=np.array([[l[1], l[2]] for l in limits]))
    env.close()
    return rec


def main():
    import pybullet  # noqa: F401  (fail early and clearly where it is missing)
    import mocca_envs  # noqa: F401  (registers the gym ids)
    os.makedirs(OUT, exist_ok=True)
    for env_id, tag in (("mocca_envs:Walker3DStepperEnv-v0", "w3d"), ("mocca_envs:MikeStepperEnv-v0", "mike")):
        for seed in range(8, 16):
            rec = record(env_id, seed)
            path = os.path.join(OUT, "pybullet_%s_seed%d.npz" % (tag, seed))
            np.savez_compressed(path, **rec)
            print(path, "steps", len(rec["rew"]), "return %.2f" % rec["rew"].sum())


if __name__ == "__main__":
    main()
