"""ctypes binding of ``libss_b200.so`` (the C-ABI in include/ss_env.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``build_native()`` with nvcc for sm_100a.
There is no CPU fallback: if the library is missing and cannot be built, or no CUDA device is visible,
the product path raises -- it never routes through ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libss_b200.so")
INCLUDE = os.path.join(REPO_DIR, "include")

NQ, ND, NJ, OBS, NWARM, NSCAL, NINT, MTW, GRID_CELLS = 28, 27, 21, 60, 16, 16, 8, 625, 11 ** 5
NEV = 18         # int32 per env and substep in ss_substep_events
TABLE_RING = 8   # sampling-table versions kept alive on the device (SS_TABLE_RING)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class StateView(ctypes.Structure):
    """``ss_state_view`` (include/ss_env.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("q", "qd", "warm", "terrain", "plank", "scalars", "ints", "mt", "table_slot")]


def _sources():
    return [os.path.join(CSRC, "ss_kernels.cu")]


def _deps():
    return _sources() + [os.path.join(CSRC, "ss_device.cuh"), os.path.join(INCLUDE, "ss_env.h"),
                         os.path.join(INCLUDE, "ss_model.h")]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_native(force=False, verbose=False):
    """Compile csrc/*.cu into libss_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("libss_b200.so is missing or stale and nvcc was not found to build it")
    # Several ranks of one torchrun job may find the library stale at the same moment: build under an exclusive file lock,
    # into a temporary file that is renamed into place (a reader never sees a half-written .so), and re-check after
    # acquiring the lock so that only the first rank compiles.
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-I", CSRC, "-o", tmp] + _sources()
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None

_P = ctypes.c_void_p
_I32 = ctypes.c_int32
_SIGNATURES = {
    "ss_create": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P, _P]),
    "ss_destroy": (None, [_P]),
    "ss_reset_all": (ctypes.c_int, [_P, _P, _P]),
    "ss_step": (ctypes.c_int, [_P] * 9),
    "ss_step_injected": (ctypes.c_int, [_P] * 9),
    "ss_step_host": (ctypes.c_int, [_P] * 8),
    "ss_reset_host": (ctypes.c_int, [_P, _P]),
    "ss_host_alloc": (ctypes.c_int, [ctypes.c_size_t, _P]),
    "ss_host_free": (None, [_P]),
    "ss_set_sample_cdf": (ctypes.c_int, [_P, _P, _I32, _I32, _P]),
    "ss_table_ring_state": (ctypes.c_int, [_P, _P, _P]),
    "ss_update_tables": (ctypes.c_int, [_P, _P, _P, _P]),
    "ss_temp_state_cells": (ctypes.c_int64, [_I32]),
    "ss_create_temp_states": (ctypes.c_int, [_P, _I32, _P, _P, _P, _P]),
    "ss_get_state": (ctypes.c_int, [_P, _P]),
    "ss_set_state": (ctypes.c_int, [_P, _P]),
    "ss_enable_substep_trace": (ctypes.c_int, [_P, _I32]),
    "ss_substep_trace": (ctypes.c_int, [_P, _P, _P]),
    "ss_substep_events": (ctypes.c_int, [_P, _P]),
    "ss_aba_substep": (ctypes.c_int, [_P, _P, _P, _P, _I32, _P]),
    "ss_gae": (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, ctypes.c_float, ctypes.c_float, _P]),
    "ss_env_post": (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    "ss_debug_floats": (_I32, []),
    "ss_enable_debug": (ctypes.c_int, [_P, _I32]),
    "ss_debug_read": (ctypes.c_int, [_P, _P]),
    "ss_host_timing": (ctypes.c_int, [_P, _P]),
    "ss_error_flags": (ctypes.c_int, [_P, _P, _I32]),
    "ss_num_envs": (_I32, [_P]),
    "ss_launch_count": (ctypes.c_int64, [_P]),
    "ss_last_error": (ctypes.c_char_p, []),
}
EXPORTS = tuple(_SIGNATURES)


def load():
    """Load (building first if stale and nvcc is present) and type the C-ABI.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("SS_B200_LIB")          # experiments only: a differently-tuned build of the same sources
    if override:
        lib = ctypes.CDLL(override)
        if hasattr(lib, "simt_emu_set_order"):
            # the host-side SIMT emulation of tests/ is test infrastructure: the product has no CPU path
            raise RuntimeError("SS_B200_LIB points at the SIMT emulation build; the product path needs the CUDA library")
    else:
        if needs_build():
            build_native()
        lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class NativeError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().ss_last_error()
        raise NativeError("%s failed (%d): %s" % (what or "libss_b200 call", rc, (msg or b"").decode()))


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
