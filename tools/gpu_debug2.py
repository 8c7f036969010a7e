import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_gpu_env, make_oracle, random_actions
np.set_printoptions(precision=5, suppress=True, linewidth=200)
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 64
a = random_actions(np.random.RandomState(100), n)
for ns in (3, 4):
    prm = dict(n_substeps=ns)
    o = make_oracle(kind, n, params=prm); g = make_gpu_env(kind, n, params=prm)
    o.reset(); g.reset()
    g.enable_substep_trace(True)
    for s in range(2):
        st = o.get_state(); g.set_state(st)
        o.step(a); g.step(a)
        qo, qdo = o.substep_trace(); qg, qdg = g.substep_trace()
        for ss in range(ns):
            ev = np.abs(qdo[ss] - qdg[ss]).max(axis=1)
            w = int(np.argmax(ev))
            print("ns=%d step %d substep %d: max dv %.3e (env %d, dof %d) max dq %.3e  bad envs %d" % (ns, s, ss, ev.max(), w, int(np.argmax(np.abs(qdo[ss][w] - qdg[ss][w]))), np.abs(qo[ss] - qg[ss]).max(), (ev > 1e-2).sum()))
        so, sg = o.get_state(), g.get_state()
        print("   after step: warm diff %.3e  ints equal %s" % (np.abs(so["warm"] - sg["warm"]).max(), np.array_equal(so["ints"], sg["ints"])))
