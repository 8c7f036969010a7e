"""GPU-box debugging aid: per-substep teacher-forced comparison with candidate / row dumps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_gpu_env, make_oracle, random_actions

np.set_printoptions(precision=5, suppress=True, linewidth=200)
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
k_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 64
prm = dict(n_substeps=1)
o = make_oracle(kind, n, params=prm); g = make_gpu_env(kind, n, params=prm)
o.reset(); g.reset()
rng = np.random.RandomState(k_steps)
for _ in range(4 * k_steps):
    o.step(random_actions(rng, n))
g.enable_debug(True); g.enable_substep_trace(True)
a = random_actions(np.random.RandomState(100 + k_steps), n)
for s in range(6):
    st = o.get_state(); g.set_state(st)
    cands = [o.contacts(i) for i in range(n)]
    o.step(a); g.step(a)
    qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
    ev = np.abs(qdo[0] - qdg[0]).max(axis=1)
    worst = int(np.argmax(ev))
    dbg = g.debug_read()
    print("substep %d: max dv %.3e (env %d)  max dq %.3e   #envs dv>1e-3: %d" % (s, ev.max(), worst, np.abs(qo[0] - qg[0]).max(), (ev > 1e-3).sum()))
    if ev.max() > 1e-3:
        i = worst
        co = cands[i]; cg = dbg[i, 756:756 + 256].reshape(32, 8)
        print(" oracle active slots", np.nonzero(co[:, 0])[0], "depth", co[co[:, 0] > 0, 1])
        print(" gpu    active slots", np.nonzero(cg[:, 0])[0], "depth", cg[cg[:, 0] > 0, 1])
        nl, nc = dbg[i, 1012], dbg[i, 1013]
        print(" gpu nl %d nc %d" % (nl, nc))
        rows = dbg[i, 1016:1016 + 192].reshape(32, 6)
        print(" gpu rows [type invD rhs rel diag lam]:\n", rows[: int(nl + 3 * nc)])
        print(" warm in state:", st["warm"][i][st["warm"][i] != 0], "at", np.nonzero(st["warm"][i])[0])
        print(" q lo/hi violations:", np.nonzero((st["q"][i, 7:] <= np.array(o.model.q_lo)) | (st["q"][i, 7:] >= np.array(o.model.q_hi)))[0])
        print(" dv oracle-gpu:", (qdo[0][i] - qdg[0][i]))
        print(" v oracle:", qdo[0][i])
        break

print("---- warm-start cache produced by each side after the teacher-forced substeps ----")
so, sg = o.get_state(), g.get_state()
dw = np.abs(so["warm"] - sg["warm"]).max(axis=1)
i = int(np.argmax(dw))
print("max |warm diff| %.3e (env %d)" % (dw.max(), i))
print(" oracle warm", so["warm"][i][so["warm"][i] != 0], np.nonzero(so["warm"][i])[0])
print(" gpu    warm", sg["warm"][i][sg["warm"][i] != 0], np.nonzero(sg["warm"][i])[0])

print("---- chained: n_substeps=2 from a common state ----")
prm2 = dict(n_substeps=2)
o2 = make_oracle(kind, n, params=prm2); g2 = make_gpu_env(kind, n, params=prm2)
o2.reset(); g2.reset()
g2.enable_substep_trace(True)
for s in range(3):
    st = o2.get_state(); g2.set_state(st)
    o2.step(a); g2.step(a)
    qo, qdo = o2.substep_trace(); qg, qdg = g2.substep_trace()
    for ss in range(2):
        ev = np.abs(qdo[ss] - qdg[ss]).max(axis=1)
        print("step %d substep %d: max dv %.3e (env %d) max dq %.3e" % (s, ss, ev.max(), int(np.argmax(ev)), np.abs(qo[ss] - qg[ss]).max()))
