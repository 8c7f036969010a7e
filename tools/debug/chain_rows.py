"""GPU box: rows / contacts of chain substep S (run with n_substeps = S + 1 so that it is the dumped one) against the same
substep teacher-forced from the oracle's state.  usage: python tools/debug/chain_rows.py kind k_steps env substep"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_gpu_env, make_oracle, random_actions
np.set_printoptions(precision=6, suppress=True, linewidth=240)
kind, k_steps, i, S = (int(x) for x in sys.argv[1:5])
n = 64
o = make_oracle(kind, n); o.reset()
rng0 = np.random.RandomState(k_steps)
for _ in range(k_steps):
    o.step(random_actions(rng0, n))
st = o.get_state()
a = random_actions(np.random.RandomState(100 + k_steps), n)
def dump(g):
    dbg = g.debug_read()[i]
    rows = dbg[756 + 256 + 4: 756 + 256 + 4 + 32 * 6].reshape(32, 6)
    con = dbg[756:756 + 128].reshape(8, 16)
    A = dbg[756 + 256 + 4 + 192:].reshape(32, 32)
    return dbg[756 + 256: 756 + 256 + 2], rows, con, dbg[:729].reshape(27, 27), dbg[729:756]
# chain: S+1 substeps
prm = dict(n_substeps=S + 1)
oc = make_oracle(kind, n, params=prm); gc = make_gpu_env(kind, n, params=prm)
oc.reset(); gc.reset(); oc.set_state(st); gc.set_state(st)
gc.enable_substep_trace(True); gc.enable_debug(True)
oc.step(a); gc.step(a)
qo, vo = oc.substep_trace(); qg, vg = gc.substep_trace()
print("chain: dv per substep", [float(np.abs(vo[s, i] - vg[s, i]).max()) for s in range(S + 1)])
nlc, rows_c, con_c, H_c, C_c = dump(gc)
# teacher-forced: oracle state after S substeps (S-substep oracle run), then one substep on both
if S > 0:
    op = make_oracle(kind, n, params=dict(n_substeps=S)); op.reset(); op.set_state(st); op.step(a); stS = op.get_state()
else:
    stS = st
# NOTE the held damping torque differs (recomputed from the current qd), so this is a neighbouring problem, not the same one
g1 = make_gpu_env(kind, n, params=dict(n_substeps=1)); g1.reset(); g1.set_state(stS); g1.enable_debug(True); g1.step(a)
nl1, rows_1, con_1, H_1, C_1 = dump(g1)
print("nl nc chain", nlc, "forced", nl1)
print("rows chain (rtype invD rhs rel diag lam):"); print(rows_c[rows_c[:, 0] >= 0])
print("rows forced:"); print(rows_1[rows_1[:, 0] >= 0])
print("contacts chain:"); print(con_c[con_c[:, 0] > 0][:, :15])
print("contacts forced:"); print(con_1[con_1[:, 0] > 0][:, :15])
print("H diff max", np.abs(H_c - H_1).max(), "C diff max", np.abs(C_c - C_1).max())
print("state diff entering the substep (gpu chain vs oracle): dq", np.abs(qo[S - 1, i] - qg[S - 1, i]).max() if S else 0.0)
print("warm gpu chain", gc.get_state()["warm"][i]); print("warm oracle chain", oc.get_state()["warm"][i])
