"""GPU box: replay a golden tape through the oracle and the GPU in lockstep; print the first step where their states differ.
usage: python tools/debug/golden_lockstep.py mike_ghost_sampleprob"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_helpers import golden_params
from test_gpu_parity import _GoldenBackend
from test_oracle_golden import _oracle_for
from steppingstone_b200 import seeding
np.set_printoptions(precision=9, suppress=True, linewidth=200)
g = np.load(os.path.join(ROOT, "tests", "golden", sys.argv[1] + ".npz"))
o = _oracle_for(g); b = _GoldenBackend(g)
if int(g["curriculum"]) >= 0:
    cdf = seeding.choice_cdf(seeding.curriculum_table(int(g["curriculum"])))
    o.set_sample_cdf(cdf); b.set_sample_cdf(cdf)
o.reset(); b.reset()
sp_step = list(g["sp_step"]) if "sp_step" in g.files else []
for t in range(len(g["rew"])):
    args = (g["actions"][t][None], g["q"][t][None], g["qd"][t][None], g["contact"][t][None])
    ro = o.step_injected(*args); rb = b.step_injected(*args)
    if t in sp_step:
        cdf = seeding.choice_cdf(g["sp_tables"][sp_step.index(t)].astype(np.float64))
        o.set_sample_cdf(cdf, apply_now=True); b.set_sample_cdf(cdf, apply_now=True)
    so, sb = o.get_state(), b.get_state()
    dt = np.abs(so["terrain"] - sb["terrain"]).max(); mt_eq = np.array_equal(so["mt"], sb["mt"]); it_eq = np.array_equal(so["ints"], sb["ints"])
    if dt > 1e-9 or not mt_eq or not it_eq or abs(float(ro[1][0]) - float(rb[1][0])) > 1e-4:
        print("step", t, "sp" if t in sp_step else "", "terrain diff", dt, "mt equal", mt_eq, "ints", so["ints"][0], sb["ints"][0], "rew", ro[1][0], rb[1][0], "ref", g["rew"][t])
        k = int(np.abs(so["terrain"][0] - sb["terrain"][0]).max(axis=1).argmax())
        print(" stone", k, "oracle", so["terrain"][0, k], "gpu", sb["terrain"][0, k])
        print(" mt pos oracle", so["mt"][0, :, 624], "gpu", sb["mt"][0, :, 624])
        break
else:
    print("no divergence over", len(g["rew"]), "steps")
