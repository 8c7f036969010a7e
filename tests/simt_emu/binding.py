"""TEST INFRASTRUCTURE ONLY: ctypes binding of tests/simt_emu/_build/libss_emu.so -- the product's CUDA sources compiled
for the host on the SIMT emulation of this directory (cuda_runtime.h, emu_runtime.cpp, build_emu.py).

``EmuEnv`` has the shape of ``oracle.binding.Oracle`` (reset / step / get_state / set_state / substep_trace) so that the
parity tests read the same.  Nothing under steppingstone_b200/ imports this; the struct definitions and the C-ABI
signature table are the product's (steppingstone_b200.model / native), the library is not.
"""
import ctypes

import numpy as np

from steppingstone_b200 import model, native, seeding

from . import build_emu

OBS, NJ = 60, 21
_libs = {}
# tag -> extra compile flags: kernel variants kept verified, and the production CTA shape (20 warps per CTA: ghost warps,
# per-substep block barrier, contact-sorted env order) next to the default one-warp-per-CTA emulation build
VARIANTS = {"": (), "sparse": ("-DSS_SPARSE_LDL=1",), "dense": ("-DSS_SPARSE_LDL=0",), "wpb20": ("-DSS_WPB=20",)}


def load(variant=""):
    if variant not in _libs:
        lib = ctypes.CDLL(build_emu.build(extra_flags=VARIANTS[variant], tag=variant))
        for name, (res, args) in native._SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib.simt_emu_collectives.restype = ctypes.c_ulonglong
        _libs[variant] = lib
    return _libs[variant]


class EmuEnv:
    def __init__(self, kind, n, seed=8, params=None, terrain_mode=None, variant=""):
        self.lib = load(variant)
        self.n = n
        self.model = model.make_model(kind)
        self.params = model.make_params(kind, params, terrain_mode)
        self.n_steps, self.rendered = self.params.n_steps, self.params.rendered
        mt_env = seeding.env_streams(seed, 0, n); mt_aux = seeding.aux_streams(kind, seed, 0, n)
        self.h = ctypes.c_void_p()
        rc = self.lib.ss_create(ctypes.byref(self.model), ctypes.byref(self.params), n, 0, mt_env.ctypes.data, mt_aux.ctypes.data,
                                ctypes.byref(self.h))
        assert rc == 0, self.lib.ss_last_error()
        self._timeout = np.zeros(n, np.uint8); self._ep_ret = np.zeros(n, np.float32); self._ep_len = np.zeros(n, np.int32)

    def _ck(self, rc, what):
        assert rc == 0, "%s: %s" % (what, self.lib.ss_last_error())

    def close(self):
        if self.h:
            self.lib.ss_destroy(self.h); self.h = ctypes.c_void_p()

    def set_sample_cdf(self, cdf, apply_now=False):
        cdf = np.ascontiguousarray(cdf, dtype=np.float64)
        self._ck(self.lib.ss_set_sample_cdf(self.h, cdf.ctypes.data, cdf.size, int(apply_now), None), "ss_set_sample_cdf")

    def reset(self):
        obs = np.zeros((self.n, OBS), np.float32)
        self._ck(self.lib.ss_reset_host(self.h, obs.ctypes.data), "ss_reset_host")
        return obs

    def step(self, act):
        a = np.ascontiguousarray(act, dtype=np.float32)
        assert a.shape == (self.n, NJ)
        obs = np.zeros((self.n, OBS), np.float32); rew = np.zeros(self.n, np.float64); done = np.zeros(self.n, np.bool_)
        self._ck(self.lib.ss_step_host(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                       self._timeout.ctypes.data, self._ep_ret.ctypes.data, self._ep_len.ctypes.data), "ss_step_host")
        return obs, rew, done, self._timeout.copy(), self._ep_ret.copy(), self._ep_len.copy()

    def get_state(self):
        st = native.alloc_state(self.n, self.n_steps, self.rendered)
        v = native.make_view(st)
        self._ck(self.lib.ss_get_state(self.h, ctypes.byref(v)), "ss_get_state")
        return st

    def set_state(self, st):
        st = {k: (None if a is None else np.ascontiguousarray(a)) for k, a in st.items()}
        v = native.make_view(st)
        self._ck(self.lib.ss_set_state(self.h, ctypes.byref(v)), "ss_set_state")

    def enable_substep_trace(self, on=True):
        self._ck(self.lib.ss_enable_substep_trace(self.h, int(on)), "ss_enable_substep_trace")

    def substep_trace(self):
        q = np.zeros((4, self.n, native.NQ)); qd = np.zeros((4, self.n, native.ND))
        self._ck(self.lib.ss_substep_trace(self.h, q.ctypes.data, qd.ctypes.data), "ss_substep_trace")
        return q, qd

    def substep_events(self):
        ev = np.zeros((4, self.n, native.NEV), np.int32)
        self._ck(self.lib.ss_substep_events(self.h, ev.ctypes.data), "ss_substep_events")
        return ev

    def step_injected(self, act, q, qd, contact):
        st = self.get_state()
        st["q"][:] = np.asarray(q, dtype=np.float64).reshape(self.n, native.NQ)
        st["qd"][:] = np.asarray(qd, dtype=np.float64).reshape(self.n, native.ND)
        st["scalars"][:, 14:16] = np.asarray(contact, dtype=np.float64).reshape(self.n, 2)
        self.set_state({"q": st["q"], "qd": st["qd"], "scalars": st["scalars"]})
        a = np.ascontiguousarray(act, dtype=np.float32).reshape(self.n, NJ)
        obs = np.zeros((self.n, OBS), np.float32); rew = np.zeros(self.n, np.float32); done = np.zeros(self.n, np.uint8)
        self._ck(self.lib.ss_step_injected(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                           self._timeout.ctypes.data, self._ep_ret.ctypes.data, self._ep_len.ctypes.data, None), "ss_step_injected")
        return obs, rew, done, self._timeout.copy(), self._ep_ret.copy(), self._ep_len.copy()

    # ---- the other kernels (host arrays stand in for device buffers) ----
    def env_post(self, act):
        a = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros((self.n, OBS), np.float32); rew = np.zeros(self.n, np.float32); done = np.zeros(self.n, np.uint8)
        self._ck(self.lib.ss_env_post(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, None), "ss_env_post")
        return obs, rew, done

    def aba_substep(self, q, qd, tau):
        q = np.ascontiguousarray(q, dtype=np.float32).copy(); qd = np.ascontiguousarray(qd, dtype=np.float32).copy()
        tau = np.ascontiguousarray(tau, dtype=np.float32)
        self._ck(self.lib.ss_aba_substep(self.h, q.ctypes.data, qd.ctypes.data, tau.ctypes.data, q.shape[0], None), "ss_aba_substep")
        return q, qd

    def create_temp_states(self, variant, grid, restore=None):
        G = int(self.lib.ss_temp_state_cells(int(variant)))
        out = np.zeros((self.n, G, OBS), np.float32)
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        r = None if restore is None else np.ascontiguousarray(restore, dtype=np.float64)
        self._ck(self.lib.ss_create_temp_states(self.h, int(variant), grid.ctypes.data, None if r is None else r.ctypes.data,
                                                out.ctypes.data, None), "ss_create_temp_states")
        return out

    def gae(self, rewards, value_preds, masks, bad_masks, gamma, lam):
        T, N = rewards.shape
        f = [np.ascontiguousarray(a, dtype=np.float32) for a in (rewards, value_preds, masks, bad_masks)]
        ret = np.zeros((T + 1, N), np.float32)
        self._ck(self.lib.ss_gae(f[0].ctypes.data, f[1].ctypes.data, f[2].ctypes.data, f[3].ctypes.data, ret.ctypes.data, T, N,
                                 float(gamma), float(lam), None), "ss_gae")
        return ret
