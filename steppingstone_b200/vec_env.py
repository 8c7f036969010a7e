"""On-device batch of stepping-stone envs behind the reference's vectorised-env protocol.

Drop-in for ``common.envs_utils.make_vec_envs`` / ``ShmemVecEnv`` (common/envs_utils.py:48-56, 486-620)
over ``mocca_envs:Walker3DStepperEnv-v0`` / ``mocca_envs:MikeStepperEnv-v0``: same constructor arguments,
same methods, same return shapes.  The N worker processes become one CUDA kernel launch per ``step``
(libss_b200.so, include/ss_env.h); per-worker gym wrappers (TimeLimit 1000, TimeLimitMask, Monitor --
envs_utils.py:25-40, 59-153) are folded into the kernel.

Two output modes:
  * ``output="numpy"`` (default, what playground/train.py expects): host arrays in / host arrays out
    through ``ss_step_host`` (pinned staging inside the library), ``infos`` is a tuple of N dicts.
  * ``output="torch"``: actions are a CUDA tensor, results are views of persistent CUDA tensors and
    ``infos`` is a ``StepInfo`` of CUDA tensors -- no host synchronisation per step.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np

from . import model as _model
from . import native, seeding

OBS_DIM, ACT_DIM = 60, 21
_NO_INFO = {}


class Box:
    """Minimal stand-in for ``gym.spaces.Box`` (gym is not a dependency of the product path)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class StepInfo:
    """Device-mode ``infos``: ``bad_transition`` [N] bool, ``done`` [N] bool, ``episode_r`` [N] f32 and
    ``episode_l`` [N] i32 (valid where ``done``) -- the only keys train.py:446-454 reads."""
    __slots__ = ("done", "bad_transition", "episode_r", "episode_l")

    def __init__(self, done, bad_transition, episode_r, episode_l):
        self.done, self.bad_transition, self.episode_r, self.episode_l = done, bad_transition, episode_r, episode_l


class BatchedStepperEnv:
    """N independent stepping-stone envs stepped by one kernel launch."""

    def __init__(self, env_id, seed, num_envs, log_dir=None, device=None, output="numpy", first_env=0,
                 terrain_mode=None, params=None):
        if env_id not in _model.ENV_IDS:
            raise ValueError("unknown env id %r (supported: %s)" % (env_id, sorted(_model.ENV_IDS)))
        if output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        self.env_id, self.seed, self.num_envs, self.output = env_id, int(seed), int(num_envs), output
        self.robot_kind = _model.ENV_IDS[env_id]
        self.first_env = int(first_env)
        self.log_dir = log_dir
        self._lib = native.load()
        self._model = _model.make_model(self.robot_kind)
        self._params = _model.make_params(self.robot_kind, params, terrain_mode)
        self.n_steps, self.rendered = self._params.n_steps, self._params.rendered
        self._max_episode_steps = self._params.max_episode_steps
        self.yaw_samples = np.array(self._params.yaw_samples)
        self.pitch_samples = np.array(self._params.pitch_samples)
        self.observation_space = Box(-np.inf, np.inf, (OBS_DIM,), np.float32)   # walker3d_envs.py:552-556
        self.action_space = Box(-1.0, 1.0, (ACT_DIM,), np.float32)              # robots.py:474-476
        if device is None:
            device = 0
        self.device_index = int(getattr(device, "index", device) or 0)
        mt_env = seeding.env_streams(self.seed, self.first_env, self.num_envs)
        mt_aux = seeding.aux_streams(self.robot_kind, self.seed, self.first_env, self.num_envs)
        self._h = ctypes.c_void_p()
        native.check(self._lib.ss_create(ctypes.byref(self._model), ctypes.byref(self._params), self.num_envs,
                                         self.device_index, mt_env.ctypes.data, mt_aux.ctypes.data,
                                         ctypes.byref(self._h)), "ss_create")
        self._t0 = time.time()
        self._pending = None
        self.curriculum = 0
        N = self.num_envs
        if output == "numpy":
            self._obs = np.zeros((N, OBS_DIM), np.float32)
            self._rew = np.zeros(N, np.float32)
            self._done = np.zeros(N, np.uint8)
            self._timeout = np.zeros(N, np.uint8)
            self._ep_ret = np.zeros(N, np.float32)
            self._ep_len = np.zeros(N, np.int32)
        else:
            import torch
            self._torch = torch
            dev = torch.device("cuda", self.device_index)
            self.device = dev
            self._obs = torch.zeros((N, OBS_DIM), dtype=torch.float32, device=dev)
            self._rew = torch.zeros(N, dtype=torch.float32, device=dev)
            self._done = torch.zeros(N, dtype=torch.uint8, device=dev)
            self._timeout = torch.zeros(N, dtype=torch.uint8, device=dev)
            self._ep_ret = torch.zeros(N, dtype=torch.float32, device=dev)
            self._ep_len = torch.zeros(N, dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------ VecEnv protocol
    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        """ShmemVecEnv.reset (envs_utils.py:542-548)."""
        if self.output == "numpy":
            native.check(self._lib.ss_reset_host(self._h, self._obs.ctypes.data), "ss_reset_host")
            return self._obs.copy()
        native.check(self._lib.ss_reset_all(self._h, self._obs.data_ptr(), self._stream()), "ss_reset_all")
        return self._obs

    def step_async(self, actions):
        assert len(actions) == self.num_envs                                    # envs_utils.py:551
        self._pending = actions

    def step_wait(self):
        actions, self._pending = self._pending, None
        if self.output == "numpy":
            a = np.ascontiguousarray(actions, dtype=np.float32)
            assert a.shape == (self.num_envs, ACT_DIM)
            assert np.isfinite(a).all()                                         # robots.py:290
            native.check(self._lib.ss_step_host(self._h, a.ctypes.data, self._obs.ctypes.data, self._rew.ctypes.data,
                                                self._done.ctypes.data, self._timeout.ctypes.data,
                                                self._ep_ret.ctypes.data, self._ep_len.ctypes.data), "ss_step_host")
            dones = self._done.astype(bool)
            t = round(time.time() - self._t0, 6)
            # one shared empty dict for the envs that did not finish (train.py:446-454 only reads keys)
            infos = [_NO_INFO] * self.num_envs
            for i in np.nonzero(dones)[0]:
                infos[i] = {}
                infos[i]["episode"] = {"r": round(float(self._ep_ret[i]), 6), "l": int(self._ep_len[i]), "t": t}
                if self._timeout[i]:
                    infos[i]["bad_transition"] = True
            return self._obs.copy(), self._rew.astype(np.float64), dones, tuple(infos)
        torch = self._torch
        a = actions
        if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()):
            a = torch.as_tensor(a, dtype=torch.float32, device=self.device).contiguous()
        assert a.shape == (self.num_envs, ACT_DIM)
        native.check(self._lib.ss_step(self._h, a.data_ptr(), self._obs.data_ptr(), self._rew.data_ptr(),
                                       self._done.data_ptr(), self._timeout.data_ptr(), self._ep_ret.data_ptr(),
                                       self._ep_len.data_ptr(), self._stream()), "ss_step")
        return self._obs, self._rew, self._done, StepInfo(self._done, self._timeout, self._ep_ret, self._ep_len)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ------------------------------------------------------------------ curriculum ops (envs_utils.py:560-590)
    def _set_table(self, p, apply_now=False):
        cdf = seeding.choice_cdf(p)
        stream = self._stream() if self.output == "torch" else None
        native.check(self._lib.ss_set_sample_cdf(self._h, cdf.ctypes.data, cdf.size, int(apply_now), stream), "ss_set_sample_cdf")

    def update_curriculum(self, curriculum):
        """walker3d_envs.py:1346-1356 / mike.py:856-866."""
        self.curriculum = min(int(curriculum), 5)
        self._set_table(seeding.curriculum_table(curriculum))

    def update_specialist(self, specialist):
        """walker3d_envs.py:604-622."""
        self._set_table(seeding.specialist_table(specialist))

    def update_sample_prob(self, probs):
        """walker3d_envs.py:1317-1320 through envs_utils.py:568-571.  The reference sends one table per env;
        train.py:269-271 broadcasts a single table, which is what the device keeps: ``probs`` may be one table
        [11,11,11,11,11] or a stack [N,...] of identical tables."""
        p = np.asarray(probs, dtype=np.float64)
        if p.size == native.GRID_CELLS * self.num_envs:
            p = p.reshape(self.num_envs, -1)
            if not np.allclose(p, p[:1]):
                raise NotImplementedError("per-env sampling tables are not supported: one table is shared by all envs")
            p = p[0]
        if p.size != native.GRID_CELLS:
            raise ValueError("sample_prob must have 11^5 cells")
        self._set_table(p, apply_now=True)

    def set_mirror(self, mirror):
        """Env-side no-op in the reference too (walker3d_envs.py:585-586)."""

    def set_env_params(self, params_dict):
        raise NotImplementedError("env parameters are compiled into the device tables: pass params= to the constructor")

    def set_robot_params(self, params_dict):
        raise NotImplementedError("robot parameters are compiled into the device tables: pass params= to the constructor")

    def create_temp_states(self):
        raise NotImplementedError("value-guided sampling (mike.py:735-753) is not built yet (SURVEY 8f item 3)")

    def get_images(self, mode="human"):
        raise NotImplementedError("rendering is out of scope (SURVEY 2, components 6b/16)")

    def get_mirror_indices(self):
        from .mirror import get_mirror_indices
        return get_mirror_indices()

    # ------------------------------------------------------------------ state access / probes
    def get_state(self):
        st = native.alloc_state(self.num_envs, self.n_steps, self.rendered)
        v = native.make_view(st)
        native.check(self._lib.ss_get_state(self._h, ctypes.byref(v)), "ss_get_state")
        return st

    def set_state(self, st):
        st = {k: (None if a is None else np.ascontiguousarray(a)) for k, a in st.items()}
        v = native.make_view(st)
        native.check(self._lib.ss_set_state(self._h, ctypes.byref(v)), "ss_set_state")

    def enable_substep_trace(self, on=True):
        native.check(self._lib.ss_enable_substep_trace(self._h, int(on)), "ss_enable_substep_trace")

    def substep_trace(self):
        q = np.zeros((4, self.num_envs, native.NQ))
        qd = np.zeros((4, self.num_envs, native.ND))
        native.check(self._lib.ss_substep_trace(self._h, q.ctypes.data, qd.ctypes.data), "ss_substep_trace")
        return q, qd

    def enable_debug(self, on=True):
        native.check(self._lib.ss_enable_debug(self._h, int(on)), "ss_enable_debug")

    def debug_read(self):
        out = np.zeros((self.num_envs, self._lib.ss_debug_floats()), np.float32)
        native.check(self._lib.ss_debug_read(self._h, out.ctypes.data), "ss_debug_read")
        return out

    def env_post(self, actions):
        """calc_state + calc_env_state on the current state, no physics (parity probe)."""
        torch = self._torch if self.output == "torch" else __import__("torch")
        dev = torch.device("cuda", self.device_index)
        a = torch.as_tensor(np.asarray(actions, dtype=np.float32), device=dev).contiguous()
        obs = torch.zeros((self.num_envs, OBS_DIM), dtype=torch.float32, device=dev)
        rew = torch.zeros(self.num_envs, dtype=torch.float32, device=dev)
        done = torch.zeros(self.num_envs, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        native.check(self._lib.ss_env_post(self._h, a.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), st), "ss_env_post")
        torch.cuda.synchronize(dev)
        return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()

    def aba_substep(self, q, qd, tau):
        """Stand-alone forward-dynamics substep on caller CUDA tensors (in place): q [n,28], qd [n,27], tau [n,21]."""
        torch = self._torch if self.output == "torch" else __import__("torch")
        st = ctypes.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
        native.check(self._lib.ss_aba_substep(self._h, q.data_ptr(), qd.data_ptr(), tau.data_ptr(), q.shape[0], st), "ss_aba_substep")

    @property
    def launch_count(self):
        return int(self._lib.ss_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ss_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_vec_envs(env_id, seed, num_processes, log_dir=None, **kwargs):
    """``common.envs_utils.make_vec_envs`` (envs_utils.py:48-56): env i is seeded ``seed + i``."""
    assert num_processes > 1
    return BatchedStepperEnv(env_id, seed, num_processes, log_dir=log_dir, **kwargs)
