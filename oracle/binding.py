"""ctypes binding of oracle/libss_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libss_oracle.so")

NQ, ND, NJ, OBS, NWARM, NSCAL, NINT, MTW = 28, 27, 21, 60, 16, 16, 8, 625


class StateView(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("q", "qd", "warm", "terrain", "plank", "scalars", "ints", "mt", "table_slot")]


def build(force=False):
    src = os.path.join(HERE, "ss_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def alloc_state(n, n_steps, rendered):
    return dict(
        q=np.zeros((n, NQ)), qd=np.zeros((n, ND)), warm=np.zeros((n, NWARM)),
        terrain=np.zeros((n, n_steps, 6)), plank=np.zeros((n, rendered, 6)),
        scalars=np.zeros((n, NSCAL)), ints=np.zeros((n, NINT), np.int32),
        mt=np.zeros((n, 2, MTW), np.uint32), table_slot=np.zeros(n, np.int32))


def make_view(state):
    v = StateView()
    for k, _ in StateView._fields_:
        a = state.get(k)
        setattr(v, k, None if a is None else a.ctypes.data)
    return v


class Oracle:
    """N envs stepped on the CPU by the C restatement."""

    def __init__(self, model, params, n, mt_env, mt_aux=None, threads=0):
        self.lib = ctypes.CDLL(build())
        self.lib.sso_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]
        self.n = n
        self.model, self.params = model, params
        self.n_steps, self.rendered = params.n_steps, params.rendered
        self.h = ctypes.c_void_p()
        mt_env = np.ascontiguousarray(mt_env, dtype=np.uint32)
        assert mt_env.shape == (n, MTW)
        g = None if mt_aux is None else np.ascontiguousarray(mt_aux, dtype=np.uint32)
        assert g is None or g.shape == (n, MTW)
        rc = self.lib.sso_create(ctypes.byref(model), ctypes.byref(params), n, _ptr(mt_env),
                                 None if g is None else _ptr(g), ctypes.byref(self.h))
        assert rc == 0, rc
        self.lib.sso_set_threads(threads)

    def close(self):
        if self.h:
            self.lib.sso_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        obs = np.zeros((self.n, OBS), np.float32)
        assert self.lib.sso_reset_all(self.h, _ptr(obs)) == 0
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        assert act.shape == (self.n, NJ)
        obs = np.zeros((self.n, OBS), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        timeout = np.zeros(self.n, np.uint8)
        ep_ret = np.zeros(self.n, np.float32)
        ep_len = np.zeros(self.n, np.int32)
        assert self.lib.sso_step(self.h, _ptr(act), _ptr(obs), _ptr(rew), _ptr(done), _ptr(timeout),
                                 _ptr(ep_ret), _ptr(ep_len)) == 0
        return obs, rew, done, timeout, ep_ret, ep_len

    def step_injected(self, act, q, qd, contact):
        """env_step with the physics replaced by the given post-physics state (golden tests)."""
        act = np.ascontiguousarray(act, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
        contact = np.ascontiguousarray(contact, dtype=np.float64)
        assert q.shape == (self.n, NQ) and qd.shape == (self.n, ND) and contact.shape == (self.n, 2)
        obs = np.zeros((self.n, OBS), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        timeout = np.zeros(self.n, np.uint8)
        ep_ret = np.zeros(self.n, np.float32)
        ep_len = np.zeros(self.n, np.int32)
        assert self.lib.sso_step_injected(self.h, _ptr(act), _ptr(q), _ptr(qd), _ptr(contact), _ptr(obs), _ptr(rew),
                                          _ptr(done), _ptr(timeout), _ptr(ep_ret), _ptr(ep_len)) == 0
        return obs, rew, done, timeout, ep_ret, ep_len

    def feet_fk(self, q):
        out = np.zeros(6)
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert self.lib.sso_feet_fk(self.h, _ptr(q), _ptr(out)) == 0
        return out.reshape(2, 3)

    def env_post(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros((self.n, OBS), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        assert self.lib.sso_env_post(self.h, _ptr(act), _ptr(obs), _ptr(rew), _ptr(done)) == 0
        return obs, rew, done

    def set_sample_cdf(self, cdf, apply_now=False):
        cdf = np.ascontiguousarray(cdf, dtype=np.float64)
        assert self.lib.sso_set_sample_cdf(self.h, _ptr(cdf), cdf.size, int(apply_now)) == 0

    def create_temp_states(self, variant, grid, restore=None):
        """create_temp_states_1 / _2 / create_temp_states (variant 1 / 2 / 5) -> [n, G, 60] float32."""
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        assert grid.shape == (5, 11)
        G = {1: 121, 2: 1331, 5: 161051}[variant]
        out = np.zeros((self.n, G, OBS), np.float32)
        r = None if restore is None else np.ascontiguousarray(restore, dtype=np.float64)
        self.lib.sso_create_temp_states.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        assert self.lib.sso_create_temp_states(self.h, variant, _ptr(grid), None if r is None else _ptr(r), _ptr(out)) == 0
        return out

    def get_state(self):
        st = alloc_state(self.n, self.n_steps, self.rendered)
        v = make_view(st)
        assert self.lib.sso_get_state(self.h, ctypes.byref(v)) == 0
        return st

    def set_state(self, st):
        st = {k: (None if a is None else np.ascontiguousarray(a)) for k, a in st.items()}
        v = make_view(st)
        assert self.lib.sso_set_state(self.h, ctypes.byref(v)) == 0

    def substep_trace(self):
        q = np.zeros((4, self.n, NQ))
        qd = np.zeros((4, self.n, ND))
        assert self.lib.sso_substep_trace(self.h, _ptr(q), _ptr(qd)) == 0
        return q, qd

    def substep_events(self):
        ev = np.zeros((4, self.n, 18), np.int32)
        assert self.lib.sso_substep_events(self.h, _ptr(ev)) == 0
        return ev

    def aba_substep(self, q, qd, tau):
        q = np.ascontiguousarray(q, dtype=np.float64).copy()
        qd = np.ascontiguousarray(qd, dtype=np.float64).copy()
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        assert self.lib.sso_aba_substep(self.h, _ptr(q), _ptr(qd), _ptr(tau), q.shape[0]) == 0
        return q, qd

    def mass_matrix(self, i=0):
        H = np.zeros((ND, ND))
        C = np.zeros(ND)
        assert self.lib.sso_mass_matrix(self.h, i, _ptr(H), _ptr(C)) == 0
        return H, C

    def invariants(self, q, qd):
        out = np.zeros(12)
        q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
        assert self.lib.sso_invariants(self.h, _ptr(q), _ptr(qd), _ptr(out)) == 0
        return dict(mass=out[0], com=out[1:4].copy(), P=out[4:7].copy(), L=out[7:10].copy(), ke=out[10], pe=out[11])

    def contacts(self, i=0):
        out = np.zeros((8, 16))
        assert self.lib.sso_contacts(self.h, i, _ptr(out)) == 0
        return out

    def gae(self, rewards, value_preds, masks, bad_masks, gamma, lam):
        T, N = rewards.shape
        ret = np.zeros((T + 1, N), np.float32)
        f = [np.ascontiguousarray(a, dtype=np.float32) for a in (rewards, value_preds, masks, bad_masks)]
        self.lib.sso_gae.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_float]
        assert self.lib.sso_gae(_ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(f[3]), _ptr(ret), T, N, gamma, lam) == 0
        return ret


def mt_doubles(words, n):
    lib = ctypes.CDLL(build())
    w = np.ascontiguousarray(words, dtype=np.uint32).copy()
    out = np.zeros(n)
    lib.sso_mt_doubles(_ptr(w), _ptr(out), n)
    return out, w
