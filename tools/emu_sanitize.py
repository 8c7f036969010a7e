#!/usr/bin/env python
"""AddressSanitizer + UndefinedBehaviorSanitizer over the product's CUDA sources on the SIMT emulation (tests/simt_emu):
out-of-bounds accesses to "device" arrays / static shared arrays, misaligned vector accesses, shift and integer overflows in
the kernels -- on the CPU, next to compute-sanitizer's memcheck on the device (profiles/r01j_sanitizer.txt).

    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emu_sanitize.py

Workload, both robots: reset, 40 control steps with falls and in-kernel auto-resets, a stone advance, create_temp_states
(2-D and 3-D grids), update_sample_prob's immediate re-sample, the stand-alone ABA kernel, the GAE kernel.
(Fiber stacks are switched by hand, so ASan's stack instrumentation is off: --param asan-stack=0.)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from simt_emu import binding
binding.VARIANTS["asan"] = ("-fsanitize=address,undefined", "--param", "asan-stack=0", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined")
from simt_emu.binding import EmuEnv
from conftest import make_oracle, random_actions
from steppingstone_b200 import seeding
for kind in (0,1):
    n=6
    g=EmuEnv(kind,n,variant="asan"); o=make_oracle(kind,n)
    if kind==1:
        cdf=seeding.choice_cdf(seeding.curriculum_table(5)); g.set_sample_cdf(cdf); o.set_sample_cdf(cdf)
    g.reset(); o.reset(); rng=np.random.RandomState(1)
    for t in range(40):
        a=random_actions(rng,n, sd=0.6); g.step(a)
    # walking-ish states incl. stone advance: teacher-force from an oracle ghost advance
    st=o.get_state(); st["ints"][:,1]=29; st["scalars"][:,14:16]=1.0
    tgt=st["terrain"][np.arange(n), st["ints"][:,0]]; st["q"][:,0:2]=tgt[:,0:2]; st["q"][:,2]=tgt[:,2]+0.9
    g.set_state(st); g.env_post(np.zeros((n,21),np.float32))
    ts=g.create_temp_states(1, seeding.fake_sample_grid(), seeding.temp_state_restore(kind))
    ts=g.create_temp_states(2, seeding.fake_sample_grid(), seeding.temp_state_restore(kind))
    rngt=np.random.RandomState(12); table=(rngt.rand(11,11,11,11,11)*(rngt.rand(11,11,11,11,11)<0.02)); table[5,5,0,5,5]+=1e-3
    g.set_sample_cdf(seeding.choice_cdf(table), apply_now=True)
    q=np.zeros((8,28)); q[:,6]=1; g.aba_substep(q, np.zeros((8,27)), np.zeros((8,21)))
    r=np.zeros((5,7),np.float32); g.gae(r, np.zeros((6,7),np.float32), np.ones((6,7),np.float32), np.ones((6,7),np.float32), 0.99, 0.95)
    g.close(); print("kind", kind, "clean")
