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
import json
import os
import sys
import time
from collections.abc import Sequence
from itertools import repeat

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


class _MonitorLog:
    """``Monitor``'s results files (common/envs_utils.py:70-94, 172-194), written when ``log_dir`` is given
    (make_env_fns, envs_utils.py:33-36): one ``<log_dir>/<rank>.monitor.csv`` per env -- a ``# {"t_start", "env_id"}`` JSON
    header line, the CSV header ``r,l,t`` and one row per finished episode.  The reference holds one open file per worker
    process; with thousands of envs in one process the rows are buffered and appended in batches (and on ``close``)."""
    EXT = "monitor.csv"

    def __init__(self, log_dir, first_env, n, env_id, t_start, flush_rows=8192):
        self.dir, self.first, self.n, self.flush_rows = log_dir, first_env, n, flush_rows
        os.makedirs(log_dir, exist_ok=True)
        header = "# {} \n".format(json.dumps({"t_start": t_start, "env_id": env_id.split(":")[-1]})) + "r,l,t\r\n"
        for i in range(n):
            with open(self._path(i), "wt") as f:
                f.write(header)
        self.rows, self.count = {}, 0

    def _path(self, i):
        return os.path.join(self.dir, "%d.%s" % (self.first + i, self.EXT))

    def write(self, idx, rs, ls, t):
        for i, r, l in zip(idx.tolist(), rs, ls):
            self.rows.setdefault(i, []).append("%s,%s,%s\r\n" % (r, l, t))     # csv.DictWriter's row format
        self.count += len(rs)
        if self.count >= self.flush_rows:
            self.flush()

    def flush(self):
        for i, rows in self.rows.items():
            with open(self._path(i), "at") as f:
                f.write("".join(rows))
        self.rows, self.count = {}, 0


class _Infos(Sequence):
    """``infos`` of a numpy-mode step: behaves like the tuple of N dicts ShmemVecEnv.step_wait returns (len, indexing,
    iteration, the keys train.py:446-454 reads).  Nothing is built until somebody looks: only the envs that finished get
    a dict of their own (``{"episode": {"r", "l", "t"}}`` from Monitor.update, envs_utils.py:136-151, plus
    ``"bad_transition"`` from TimeLimitMask, :62-63), every other position is one shared empty dict."""
    __slots__ = ("_n", "_raw", "_own")

    def __init__(self, n, raw):
        self._n, self._raw, self._own = n, raw, None

    def _dicts(self):
        if self._own is None:
            own = {}
            if self._raw is not None:
                idx, rs, ls, tos, t = self._raw
                for i, r, l, to in zip(idx, rs, ls, tos):
                    info = {"episode": {"r": r, "l": l, "t": t}}
                    if to:
                        info["bad_transition"] = True
                    own[i] = info
            self._own = own
        return self._own

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return tuple(self[k] for k in range(*i.indices(self._n)))
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._dicts().get(i, _NO_INFO)

    def __iter__(self):
        own = self._dicts()
        if not own:
            return repeat(_NO_INFO, self._n)
        return (own.get(i, _NO_INFO) for i in range(self._n))


class _ArrayPool:
    """Output arrays of the numpy-mode step.  The reference returns FRESH arrays every step (envs_utils.py:556-558,
    608-620) and callers may keep them, so an array is handed out again only when nobody else holds a reference to it
    (refcount check); otherwise a new one is allocated.  Why bother: a fresh 2 MB ``np.empty`` is an mmap whose pages
    fault in while the C side unpacks into it and an munmap when it dies -- ~100 us per step at 8192 envs."""

    def __init__(self, shape, dtype, depth=4, factory=None):
        self.shape, self.dtype, self.depth, self.items = shape, dtype, depth, []
        self.factory = factory or (lambda: np.empty(shape, dtype))

    def get(self):
        """-> (array, its data pointer): ``ndarray.ctypes.data`` builds a ctypes object on every access (~1 us each)."""
        for a, ptr in self.items:
            if sys.getrefcount(a) == 3:           # the tuple's reference, the loop variable, getrefcount's argument
                return a, ptr
        a = self.factory()
        ptr = a.ctypes.data
        if len(self.items) < self.depth:
            self.items.append((a, ptr))
        return a, ptr

    def prefill(self, n=2):
        """Allocate the arrays a caller that keeps the previous step's outputs alternates between (pinning memory costs
        milliseconds: not inside somebody's first steps)."""
        while len(self.items) < min(n, self.depth):
            a = self.factory()
            self.items.append((a, a.ctypes.data))


class _PinnedBlock:
    """cudaHostAlloc'ed memory owned by a Python object (freed when the last array over it dies)."""

    def __init__(self, lib, nbytes):
        self.lib = lib
        p = ctypes.c_void_p()
        native.check(lib.ss_host_alloc(nbytes, ctypes.byref(p)), "ss_host_alloc")
        self.ptr = p.value

    def __del__(self):
        try:
            self.lib.ss_host_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def _pinned_array(lib, shape, dtype):
    """A numpy array over pinned host memory: ``ss_step_host`` lets the kernel write the observations straight into it
    (the role ShmemVecEnv's shared-memory observation buffers play in the reference, envs_utils.py:507-513)."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    blk = _PinnedBlock(lib, nbytes)
    buf = (ctypes.c_char * nbytes).from_address(blk.ptr)
    buf._blk = blk                              # array -> memoryview -> ctypes buffer -> block
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


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
        self._tables = _model.load_tables(self.robot_kind)
        self._overrides = dict(params or {})
        self._terrain_mode = terrain_mode
        self._power = 1.0
        self._model = _model.make_model(self.robot_kind, self._tables)
        self._params = _model.make_params(self.robot_kind, self._overrides, terrain_mode)
        self._fake_grid = seeding.fake_sample_grid()
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
        self._no_infos = (_NO_INFO,) * self.num_envs
        self.curriculum = 0
        self._table = None                                                  # uniform until a table is set
        self._table_log = []                                                # the table versions set so far, oldest first (<= TABLE_RING)
        N = self.num_envs
        if log_dir and output == "torch":
            raise NotImplementedError("log_dir (Monitor csv files) needs output='numpy': device mode keeps the episode "
                                      "statistics on the device (steppingstone_b200.collector.EpisodeStats)")
        self._monitor = _MonitorLog(log_dir, self.first_env, N, env_id, self._t0) if log_dir else None
        if output == "numpy":
            # every output array is pinned host memory: ss_step_host lets the kernel write into it directly
            self._timeout = _pinned_array(self._lib, (N,), np.uint8); self._timeout[:] = 0
            self._ep_ret = _pinned_array(self._lib, (N,), np.float32); self._ep_ret[:] = 0
            self._ep_len = _pinned_array(self._lib, (N,), np.int32); self._ep_len[:] = 0
            self._stat_ptrs = (self._timeout.ctypes.data, self._ep_ret.ctypes.data, self._ep_len.ctypes.data)
            self._pool_obs = _ArrayPool((N, OBS_DIM), np.float32, factory=lambda: _pinned_array(self._lib, (N, OBS_DIM), np.float32))
            self._pool_rew = _ArrayPool((N,), np.float64, factory=lambda: _pinned_array(self._lib, (N,), np.float64))
            self._pool_done = _ArrayPool((N,), np.bool_, factory=lambda: _pinned_array(self._lib, (N,), np.bool_))
            for pool in (self._pool_obs, self._pool_rew, self._pool_done):
                pool.prefill()
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
            obs, p_obs = self._pool_obs.get()
            native.check(self._lib.ss_reset_host(self._h, p_obs), "ss_reset_host")
            return obs
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
            N = self.num_envs
            # arrays nobody else refers to: fresh ones as far as the caller can tell (ShmemVecEnv._decode_obses / step_wait
            # return new arrays every step, envs_utils.py:556-558, 608-620)
            (obs, p_obs), (rew, p_rew), (dones, p_done) = self._pool_obs.get(), self._pool_rew.get(), self._pool_done.get()
            rc = self._lib.ss_step_host(self._h, a.ctypes.data, p_obs, p_rew, p_done, *self._stat_ptrs)
            if rc != 0:
                assert np.isfinite(a).all()                                     # robots.py:290
                native.check(rc, "ss_step_host")
            # one shared empty dict for the envs that did not finish (train.py:446-454 only reads keys)
            idx = np.flatnonzero(dones)
            raw = None
            if idx.size:                                                        # Monitor.update, envs_utils.py:136-151
                t = round(time.time() - self._t0, 6)
                rs = np.round(self._ep_ret[idx].astype(np.float64), 6).tolist()
                ls = self._ep_len[idx].tolist()
                raw = (idx.tolist(), rs, ls, self._timeout[idx].tolist(), t)
                if self._monitor is not None:
                    self._monitor.write(idx, rs, ls, t)
            return obs, rew, dones, _Infos(N, raw)
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

    def check_errors(self, clear=True):
        """Device mode cannot raise from inside a stream-ordered ``ss_step``: a non-finite action (robots.py:290 asserts on
        it in the reference) is stepped as zero and remembered.  Call once per rollout; raises like the reference would."""
        flags = ctypes.c_int32(0)
        native.check(self._lib.ss_error_flags(self._h, ctypes.byref(flags), int(clear)), "ss_error_flags")
        assert not (flags.value & 1), "a non-finite action reached step() since the last check (robots.py:290)"

    # ------------------------------------------------------------------ curriculum ops (envs_utils.py:560-590)
    def _set_table(self, p, apply_now=False):
        self._table = np.array(p, dtype=np.float64).reshape(-1)             # kept for save_state
        self._table_log.append(self._table)                                  # the device keeps a ring of the last TABLE_RING versions
        del self._table_log[:-native.TABLE_RING]
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
        train.py:269-271 broadcasts a single table, which is what the device keeps: ``probs`` may be one table or a
        stack [N, ...] of identical tables.  Accepted table shapes: the env's 5-D grid [11,11,11,11,11]
        (yaw, pitch, r, x_tilt, y_tilt), and the 2-D [11,11] / 3-D [11,11,11] tables train.py:267-271, 311-315 build
        (``update_sample_prob_1/_2``, walker3d_envs.py:1307-1315): those are embedded at index 0 of the missing axes --
        every r / tilt sample of the stepper envs is the same value (mike.py:74-76), so this is the draw they mean.

        As in the reference (``if self.update_terrain:``, mike.py:828), ONLY the envs that advanced a stone in their last step
        take the table and re-sample their next-next stone at once; every other env keeps drawing from the table it has.
        The device holds a ring of the last 8 table versions and a version index per env (include/ss_env.h); an env that
        has not advanced a stone across 8 consecutive updates is moved to the oldest surviving version."""
        p = expand_sample_table(probs, self.num_envs)
        self._set_table(p, apply_now=True)

    def set_mirror(self, mirror):
        """Env-side no-op in the reference too (walker3d_envs.py:585-586)."""

    # reference attribute -> ss_params field (walker3d_envs.py:487-547 / mike.py:30-90)
    _ENV_PARAMS = {"done_height": "done_height", "step_radius": "step_radius", "stop_frames": "stop_frames",
                   "electricity_cost": "electricity_cost", "stall_torque_cost": "stall_torque_cost",
                   "joints_at_limit_cost": "joints_at_limit_cost", "tilt_limit": "tilt_limit_deg",
                   "yaw_samples": "yaw_samples", "pitch_samples": "pitch_samples", "r_samples": "r_samples",
                   "x_tilt_samples": "xtilt_samples", "y_tilt_samples": "ytilt_samples"}
    _ENV_FAKE = ("fake_yaw_samples", "fake_pitch_samples", "fake_r_samples", "fake_x_tilt_samples", "fake_y_tilt_samples")
    _ENV_FIXED = ("n_steps", "rendered_step_count", "lookahead", "sample_size", "yaw_sample_size", "pitch_sample_size",
                  "r_sample_size", "x_tilt_sample_size", "y_tilt_sample_size", "control_step", "llc_frame_skip",
                  "sim_frame_skip")

    def _push_tables(self):
        self._model = _model.make_model(self.robot_kind, self._tables)
        self._params = _model.make_params(self.robot_kind, self._overrides, self._terrain_mode)
        stream = self._stream() if self.output == "torch" else None
        native.check(self._lib.ss_update_tables(self._h, ctypes.byref(self._model), ctypes.byref(self._params), stream),
                     "ss_update_tables")
        self.yaw_samples = np.array(self._params.yaw_samples)
        self.pitch_samples = np.array(self._params.pitch_samples)

    def set_env_params(self, params_dict):
        """EnvBase.set_env_params (env_base.py:73-76) on every env: ``setattr`` for attributes the env has, silently
        nothing otherwise.  Attributes that size device arrays cannot change after construction and raise."""
        for k, v in params_dict.items():
            if k in self._ENV_PARAMS:
                self._overrides[self._ENV_PARAMS[k]] = (np.asarray(v, dtype=np.float64).tolist() if np.ndim(v) else v)
            elif k == "r_range":
                self._overrides["r_max"] = float(np.asarray(v).reshape(-1)[1])
            elif k in self._ENV_FAKE:
                self._fake_grid[self._ENV_FAKE.index(k)] = np.asarray(v, dtype=np.float64)
            elif k in self._ENV_FIXED:
                raise NotImplementedError("%s is fixed at construction (it sizes the device arrays): pass params= to the constructor" % k)
        self._push_tables()

    def set_robot_params(self, params_dict):
        """EnvBase.set_robot_params (env_base.py:78-85): only ``power`` matters -- the torque limits are recomputed as
        power * power_coef (robots.py:385-388)."""
        if "power" in params_dict:
            base = _model.load_tables(self.robot_kind)["tau_max"]
            self._power = float(params_dict["power"])
            self._tables["tau_max"] = base * self._power
        for ref_name, table in (("P", "pd_p"), ("D", "pd_d")):                  # robots.py:482-483 (PD actuator modes)
            if ref_name in params_dict:
                v = np.asarray(params_dict[ref_name], dtype=np.float64)
                assert v.shape == (ACT_DIM,)
                self._tables[table] = v
        self._push_tables()

    def create_temp_states(self, variant=5):
        """ShmemVecEnv.create_temp_states (envs_utils.py:573-578) -> [N, G, 60] float32: for every env that advanced a
        stone in the last step, ``get_temp_state()`` with the next-next stone moved to each cell of the probe grid
        (mike.py:735-753; ``variant`` 5 = yaw x pitch x r x x_tilt x y_tilt, 161051 cells; 2 = yaw x pitch x r
        (create_temp_states_2, 1331); 1 = yaw x pitch at r = 0.7 (create_temp_states_1, 121 -- the shape
        train.py:255-257 views).  Envs that did not advance return zeros (the reference's ``self.temp_states``)."""
        G = int(self._lib.ss_temp_state_cells(int(variant)))
        if G == 0:
            raise ValueError("variant must be 1, 2 or 5")
        nbytes = self.num_envs * G * OBS_DIM * 4
        torch = self._torch if self.output == "torch" else __import__("torch")
        dev = torch.device("cuda", self.device_index)
        free, _ = torch.cuda.mem_get_info(dev)
        if nbytes > 0.9 * free:
            raise MemoryError("create_temp_states(variant=%d) needs %.1f GB for %d envs x %d cells; use variant=1/2 or "
                              "fewer envs (the reference only calls the 5-D form on a single evaluation env)"
                              % (variant, nbytes / 1e9, self.num_envs, G))
        out = torch.empty((self.num_envs, G, OBS_DIM), dtype=torch.float32, device=dev)
        restore = seeding.temp_state_restore(self.robot_kind)
        if self.output == "torch":
            stream = self._stream()
        else:
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        native.check(self._lib.ss_create_temp_states(self._h, int(variant), self._fake_grid.ctypes.data,
                                                     None if restore is None else restore.ctypes.data,
                                                     out.data_ptr(), stream), "ss_create_temp_states")
        if self.output == "torch":
            return out
        return out.cpu().numpy()

    def create_temp_states_1(self):
        return self.create_temp_states(1)

    def create_temp_states_2(self):
        return self.create_temp_states(2)

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

    def _state_spec(self):
        N = self.num_envs
        return {"q": (np.float64, (N, native.NQ)), "qd": (np.float64, (N, native.ND)), "warm": (np.float64, (N, native.NWARM)),
                "terrain": (np.float64, (N, self.n_steps, 6)), "plank": (np.float64, (N, self.rendered, 6)),
                "scalars": (np.float64, (N, native.NSCAL)), "ints": (np.int32, (N, native.NINT)),
                "mt": (np.uint32, (N, 2, native.MTW)), "table_slot": (np.int32, (N,))}

    def set_state(self, st):
        """Any subset of the ``get_state`` arrays.  Every array is converted to the dtype the C-ABI copies (a float32 ``q``
        used to be memcpy'd as doubles) and its shape is checked."""
        spec = self._state_spec()
        out = {}
        for k, a in st.items():
            if a is None:
                continue
            if k not in spec:
                raise KeyError("unknown state array %r (expected a subset of %s)" % (k, sorted(spec)))
            dtype, shape = spec[k]
            a = np.ascontiguousarray(a, dtype=dtype)
            if a.shape != shape:
                raise ValueError("state array %r has shape %s, expected %s" % (k, a.shape, shape))
            out[k] = a
        v = native.make_view(out)
        native.check(self._lib.ss_set_state(self._h, ctypes.byref(v)), "ss_set_state")

    @staticmethod
    def _npz(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"

    def save_state(self, path):
        """Checkpoint the whole simulator to an .npz: poses, velocities, contact cache, terrain, planks, bookkeeping and
        both MT19937 streams of every env, plus what lives outside those arrays -- the sampling table in force
        (curriculum / specialist / update_sample_prob), the curriculum level, the env-parameter overrides and the robot's
        power / PD tables.  ``load_state`` on an env built with the same (env_id, seed, num_envs, first_env) continues the
        rollout bit for bit.  The reference has no equivalent (its env state dies with the worker processes)."""
        import json
        st = self.get_state()
        head, count = ctypes.c_int32(0), ctypes.c_int64(0)
        native.check(self._lib.ss_table_ring_state(self._h, ctypes.byref(head), ctypes.byref(count)), "ss_table_ring_state")
        np.savez_compressed(self._npz(path), env_id=self.env_id, seed=self.seed, first_env=self.first_env,
                            n_steps=self.n_steps, rendered=self.rendered, curriculum=self.curriculum,
                            has_table=self._table is not None, table=(self._table if self._table is not None else np.zeros(0)),
                            table_log=(np.stack(self._table_log) if self._table_log else np.zeros((0, native.GRID_CELLS))),
                            table_head=head.value, table_count=count.value,
                            overrides=json.dumps(self._overrides), terrain_mode=json.dumps(self._terrain_mode),
                            power=self._power, tau_max=self._tables["tau_max"], pd_p=self._tables["pd_p"],
                            pd_d=self._tables["pd_d"], fake_grid=self._fake_grid, **st)

    def load_state(self, path):
        import json
        with np.load(self._npz(path)) as z:
            for key, mine in (("env_id", self.env_id), ("seed", self.seed), ("first_env", self.first_env),
                              ("n_steps", self.n_steps), ("rendered", self.rendered)):
                if key in z.files and str(z[key]) != str(mine):
                    raise ValueError("checkpoint was written with %s = %s, this env has %s" % (key, z[key], mine))
            if z["q"].shape[0] != self.num_envs:
                raise ValueError("checkpoint is for %d envs, this env has %d" % (z["q"].shape[0], self.num_envs))
            if "overrides" in z.files:                                     # (older checkpoints hold the state arrays only)
                self._overrides = json.loads(str(z["overrides"]))
                self._terrain_mode = json.loads(str(z["terrain_mode"]))
                self._power = float(z["power"])
                for k in ("tau_max", "pd_p", "pd_d"):
                    self._tables[k] = np.array(z[k])
                self._fake_grid = np.array(z["fake_grid"])
                self._push_tables()
                self._max_episode_steps = self._params.max_episode_steps
                self.curriculum = int(z["curriculum"])
            slots = None
            if "table_log" in z.files:
                # The sampling-table versions the envs draw from (a ring on the device, a version index per env): re-install
                # the versions that were alive, oldest first -- the uniform table of ss_create too while it was still in the
                # ring -- then point every env at the version of the same AGE it had.
                K = native.TABLE_RING
                log = [np.array(t) for t in z["table_log"]]
                head_s, count_s = int(z["table_head"]), int(z["table_count"])
                stream = self._stream() if self.output == "torch" else None
                if count_s <= K and (log or self._table_log):
                    cdf = (np.arange(native.GRID_CELLS, dtype=np.float64) + 1.0) / native.GRID_CELLS
                    native.check(self._lib.ss_set_sample_cdf(self._h, cdf.ctypes.data, cdf.size, 0, stream), "ss_set_sample_cdf")
                self._table_log = []
                for t in log:
                    self._set_table(t)
                if not log:
                    self._table = None
                head, count = ctypes.c_int32(0), ctypes.c_int64(0)
                native.check(self._lib.ss_table_ring_state(self._h, ctypes.byref(head), ctypes.byref(count)), "ss_table_ring_state")
                age = (head_s - np.array(z["table_slot"], dtype=np.int64)) % K
                slots = ((head.value - age) % K).astype(np.int32)
            elif "overrides" in z.files and bool(z["has_table"]):
                self._set_table(np.array(z["table"]))
            state = {k: z[k] for k in ("q", "qd", "warm", "terrain", "plank", "scalars", "ints", "mt")}
            if slots is not None:
                state["table_slot"] = slots
            self.set_state(state)

    def enable_substep_trace(self, on=True):
        native.check(self._lib.ss_enable_substep_trace(self._h, int(on)), "ss_enable_substep_trace")

    def substep_trace(self):
        q = np.zeros((4, self.num_envs, native.NQ))
        qd = np.zeros((4, self.num_envs, native.ND))
        native.check(self._lib.ss_substep_trace(self._h, q.ctypes.data, qd.ctypes.data), "ss_substep_trace")
        return q, qd

    def substep_events(self):
        """[4, N, 18] int32: the discrete decisions of each substep of the last step (include/ss_env.h ss_substep_events)."""
        ev = np.zeros((4, self.num_envs, native.NEV), np.int32)
        native.check(self._lib.ss_substep_events(self._h, ev.ctypes.data), "ss_substep_events")
        return ev

    def enable_debug(self, on=True):
        native.check(self._lib.ss_enable_debug(self._h, int(on)), "ss_enable_debug")

    def debug_read(self):
        out = np.zeros((self.num_envs, self._lib.ss_debug_floats()), np.float32)
        native.check(self._lib.ss_debug_read(self._h, out.ctypes.data), "ss_debug_read")
        return out

    def step_injected(self, actions, q, qd, contact):
        """One step() with the physics replaced by a post-physics state (q [N,28], qd [N,27], foot-contact flags [N,2]):
        everything after ``stepSimulation`` runs -- calc_state, calc_env_state, rewards, TimeLimit / Monitor, auto-reset.
        Golden-tape probe (the reference-generated fixtures replay through it); returns numpy
        (obs, rew, done, timeout, ep_ret, ep_len) like the oracle's ``step_injected``."""
        torch = self._torch if self.output == "torch" else __import__("torch")
        dev = torch.device("cuda", self.device_index)
        N = self.num_envs
        st = self.get_state()
        st["q"][:] = np.asarray(q, dtype=np.float64).reshape(N, native.NQ)
        st["qd"][:] = np.asarray(qd, dtype=np.float64).reshape(N, native.ND)
        st["scalars"][:, 14:16] = np.asarray(contact, dtype=np.float64).reshape(N, 2)
        self.set_state({"q": st["q"], "qd": st["qd"], "scalars": st["scalars"]})
        a = torch.as_tensor(np.asarray(actions, dtype=np.float32).reshape(N, ACT_DIM), device=dev).contiguous()
        obs = torch.zeros((N, OBS_DIM), dtype=torch.float32, device=dev)
        rew = torch.zeros(N, dtype=torch.float32, device=dev)
        done = torch.zeros(N, dtype=torch.uint8, device=dev); to = torch.zeros(N, dtype=torch.uint8, device=dev)
        er = torch.zeros(N, dtype=torch.float32, device=dev); el = torch.zeros(N, dtype=torch.int32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        native.check(self._lib.ss_step_injected(self._h, a.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                                to.data_ptr(), er.data_ptr(), el.data_ptr(), stream), "ss_step_injected")
        torch.cuda.synchronize(dev)
        return (obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), to.cpu().numpy(), er.cpu().numpy(), el.cpu().numpy())

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
        if getattr(self, "_monitor", None) is not None:
            self._monitor.flush()
        if getattr(self, "_h", None):
            self._lib.ss_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def expand_sample_table(probs, num_envs):
    """One [11^5] sampling table from what callers pass to ``update_sample_prob`` (see there)."""
    g = native.GRID_CELLS
    p = np.asarray(probs, dtype=np.float64)
    per_env = None
    for cells in (11 ** 5, 11 ** 3, 11 ** 2):
        if p.size == cells:
            per_env = p.reshape(1, -1); break
        if p.size == cells * num_envs and num_envs > 1:
            per_env = p.reshape(num_envs, -1); break
    if per_env is None:
        raise ValueError("sample_prob must be a [11,11], [11,11,11] or [11,11,11,11,11] table (or one per env)")
    if not np.allclose(per_env, per_env[:1]):
        raise NotImplementedError("N DIFFERENT sampling tables in one update_sample_prob call are not supported (train.py:266-271 "
                                  "always sends N copies of one table).  What is per env is which VERSION an env draws from: like in "
                                  "the reference, only the envs that advanced a stone in their last step take a new table")
    t = per_env[0]
    if t.size == g:
        return t
    full = np.zeros((11,) * 5)
    if t.size == 11 ** 2:
        full[:, :, 0, 0, 0] = t.reshape(11, 11)
    else:
        full[:, :, :, 0, 0] = t.reshape(11, 11, 11)
    return full.reshape(-1)


def make_vec_envs(env_id, seed, num_processes, log_dir=None, **kwargs):
    """``common.envs_utils.make_vec_envs`` (envs_utils.py:48-56): env i is seeded ``seed + i``."""
    assert num_processes > 1
    return BatchedStepperEnv(env_id, seed, num_processes, log_dir=log_dir, **kwargs)
