"""GPU box: find the envs of test_teacher_forced_step that break the strict chain criterion and dump, substep by substep
(teacher-forced from the oracle's state), contacts and rows of both sides.  usage: python tools/debug/chain_fail.py kind k_steps"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_gpu_env, make_oracle, random_actions
from parity_helpers import TOL_Q, TOL_V
np.set_printoptions(precision=6, suppress=True, linewidth=220)
kind, k_steps = int(sys.argv[1]), int(sys.argv[2])
n = 64
o = make_oracle(kind, n); g = make_gpu_env(kind, n)
o.reset(); g.reset()
rng0 = np.random.RandomState(k_steps)
for _ in range(k_steps):
    o.step(random_actions(rng0, n))
st = o.get_state(); g.set_state(st)
g.enable_substep_trace(True)
a = random_actions(np.random.RandomState(100 + k_steps), n)
o.step(a); g.step(a)
qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
evo, evg = o.substep_events(), g.substep_events()
same = np.logical_and.accumulate((evo == evg).all(axis=2), axis=0)
clamped = np.logical_or.accumulate(((evo[:, :, 1] | evg[:, :, 1]) & (1 << 30)) != 0, axis=0)
vs = np.maximum(1.0, np.abs(qdo).max(axis=2))
e_q = np.abs(qo - qg).max(axis=2); e_v = np.abs(qdo - qdg).max(axis=2)
bad = same & ~clamped & ~((e_q < TOL_Q[:, None] * vs) & (e_v < TOL_V[:, None] * vs))
print("lib", os.environ.get("SS_B200_LIB", "default"), "bad (substep, env):", np.argwhere(bad).tolist())
for i in sorted(set(np.argwhere(bad)[:, 1].tolist())):
    print("=== env", i, "e_q", e_q[:, i], "e_v", e_v[:, i], "vscale", vs[:, i])
    print("events oracle", evo[:, i].tolist()); print("events gpu   ", evg[:, i].tolist())
    o1 = make_oracle(kind, 1, params=dict(n_substeps=1)); g1 = make_gpu_env(kind, 1, params=dict(n_substeps=1))
    o1.reset(); g1.reset()
    o1.set_state({k: v[i:i + 1].copy() for k, v in st.items()})
    g1.enable_substep_trace(True); g1.enable_debug(True)
    for s in range(4):
        stc = o1.get_state(); g1.set_state(stc)
        o1.step(a[i:i + 1]); g1.step(a[i:i + 1])
        q1, v1 = o1.substep_trace(); q2, v2 = g1.substep_trace()
        dbg = g1.debug_read()[0]
        print("-- teacher-forced substep", s, "dq", np.abs(q1[0] - q2[0]).max(), "dv", np.abs(v1[0] - v2[0]).max(), "argmax dof", int(np.abs(v1[0] - v2[0]).argmax()),
              "ev equal", bool((o1.substep_events()[0] == g1.substep_events()[0]).all()))
        if np.abs(v1[0] - v2[0]).max() > 1e-3 * max(1.0, np.abs(v1[0]).max()):
            rows = dbg[756 + 256 + 4: 756 + 256 + 4 + 32 * 6].reshape(32, 6)
            print("gpu rows (rtype invD rhs rel diag lam):"); print(rows[rows[:, 0] >= 0])
            print("gpu contacts (active depth n3 x3 id ba bb xb3):"); cg = dbg[756:756 + 128].reshape(8, 16); print(cg[cg[:, 0] > 0][:, :14])
            print("oracle state q - lo:", stc["q"][0, 7:] - np.array(o1.model.q_lo)); print("hi - q:", np.array(o1.model.q_hi) - stc["q"][0, 7:])
            print("dv per dof:", (v1[0] - v2[0])[0])
            print("oracle qd after:", v1[0][0])
            # the oracle's contact list for the SAME state: re-detect by stepping a copy with zero dt is not possible; print warm ids instead
            print("oracle warm after:", o1.get_state()["warm"][0]); print("gpu warm after:   ", g1.get_state()["warm"][0])
