#!/usr/bin/env python
"""Data-only reader for the reference's shipped policies (playground/models/*.pt) -> tests/golden/policy_*.npz.

The files are whole pickled ``nn.Module`` objects in torch's legacy (pre-zip) ``torch.save`` layout.  ``torch.load``
would import and call whatever the pickle names; this reader does NOT: it walks the pickle stream with an Unpickler
whose ``find_class`` knows a fixed allow-list of names and maps each to an INERT stand-in defined in this file
(nothing named in the checkpoint is ever imported, constructed or called), refuses every other global, and rebuilds
tensors itself as numpy arrays from the raw storage bytes.  The result is the actor's six weight matrices and biases
(``common/controller.py:SoftsignActor``: fc1..fc5, out) plus the log-std bias -- numbers, no code.

Legacy layout (torch/serialization.py, ``_legacy_load``): pickle(magic) pickle(protocol) pickle(sys_info) pickle(object,
storages by persistent id ('storage', type, key, location, numel, view)) pickle([storage keys]) then, per key, an int64
element count followed by the raw little-endian elements.

    python tools/read_legacy_checkpoint.py            # writes tests/golden/policies/{w3d_latest,w3d_base,mike_latest}.npz
"""
import collections
import io
import os
import pickle
import struct
import sys

import numpy as np

MAGIC = 0x1950a86a20f9469cfc6c
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = "/root/reference/playground/models"


class Inert:
    """Stand-in for every module class the checkpoint names: holds the attributes the pickle gives it, runs nothing."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class StorageType:
    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)


class TensorRef:
    """A tensor as the pickle describes it: storage key + offset/size/stride; resolved once the raw bytes are read."""

    def __init__(self, storage, offset, size, stride):
        self.storage, self.offset, self.size, self.stride = storage, int(offset), tuple(size), tuple(stride)

    def numpy(self, raw):
        base = raw[self.storage[0]]
        if len(self.size) == 0:
            return base[self.offset].copy()
        isz = base.itemsize
        return np.lib.stride_tricks.as_strided(base[self.offset:], self.size, tuple(s * isz for s in self.stride)).copy()


def _rebuild_tensor_v2(storage, offset, size, stride, requires_grad=False, backward_hooks=None):
    return TensorRef(storage, offset, size, stride)


def _rebuild_parameter(data, requires_grad=False, backward_hooks=None):
    return data


ALLOWED = {
    ("collections", "OrderedDict"): collections.OrderedDict,
    ("torch._utils", "_rebuild_tensor_v2"): _rebuild_tensor_v2,
    ("torch._utils", "_rebuild_parameter"): _rebuild_parameter,
    ("torch", "FloatStorage"): StorageType("<f4"),
    ("torch", "LongStorage"): StorageType("<i8"),
    ("torch.nn.backends.thnn", "_get_thnn_function_backend"): (lambda *a: None),
}
INERT_MODULES = ("common.controller", "torch.nn.modules.")     # module classes -> one fresh Inert subclass per name


class DataOnlyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in ALLOWED:
            return ALLOWED[(module, name)]
        if module == INERT_MODULES[0] or module.startswith(INERT_MODULES[1]):
            return type(name, (Inert,), {"_ref_module": module})
        raise pickle.UnpicklingError("refused global %s.%s (not on the allow-list)" % (module, name))

    def persistent_load(self, pid):
        if isinstance(pid, tuple) and len(pid) == 4 and pid[0] == "module" and isinstance(pid[1], type) and issubclass(pid[1], Inert):
            return pid[1]                # ('module', class, source file, source text): torch's source-change check; the text is dropped
        if not (isinstance(pid, tuple) and pid and pid[0] == "storage"):
            raise pickle.UnpicklingError("refused persistent id %r" % (pid,))
        _, stype, key, _location, numel, view = pid
        if not isinstance(stype, StorageType) or view is not None:
            raise pickle.UnpicklingError("refused storage record %r" % (pid,))
        return (key, stype.dtype, int(numel))


def read_checkpoint(path):
    """-> (root Inert object, {storage key: 1-D numpy array})."""
    with open(path, "rb") as fh:
        f = io.BytesIO(fh.read())
    if DataOnlyUnpickler(f).load() != MAGIC:
        raise ValueError("%s: not a legacy torch.save file" % path)
    DataOnlyUnpickler(f).load()          # protocol version
    DataOnlyUnpickler(f).load()          # sys info
    storages = {}

    class U(DataOnlyUnpickler):
        def persistent_load(self, pid):
            rec = DataOnlyUnpickler.persistent_load(self, pid)
            if isinstance(rec, tuple):
                storages[rec[0]] = rec
            return rec

    root = U(f).load()
    keys = DataOnlyUnpickler(f).load()
    raw = {}
    for key in keys:
        _, dtype, numel = storages[key]
        (n,) = struct.unpack("<q", f.read(8))
        if n != numel:
            raise ValueError("storage %s: %d elements on disk, %d in the pickle" % (key, n, numel))
        raw[key] = np.frombuffer(f.read(n * dtype.itemsize), dtype=dtype)
    return root, raw


def actor_arrays(path):
    """The deterministic policy of common/controller.py: Policy.actor (SoftsignActor fc1..fc5, out) + dist.logstd._bias."""
    root, raw = read_checkpoint(path)
    if isinstance(root, (tuple, list)):       # some checkpoints are (actor_critic, ob_rms)-style tuples
        root = root[0]
    actor = root._modules["actor"]
    out = {}
    for name in ("fc1", "fc2", "fc3", "fc4", "fc5", "out"):
        lin = actor._modules[name]
        out[name + "_w"] = lin._parameters["weight"].numpy(raw).astype(np.float32)
        out[name + "_b"] = lin._parameters["bias"].numpy(raw).astype(np.float32)
    out["actor_class"] = np.array(type(actor).__name__)
    out["logstd"] = root._modules["dist"]._modules["logstd"]._parameters["_bias"].numpy(raw).astype(np.float32).reshape(-1)
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden", "policies"), exist_ok=True)
    # Walker3DStepperEnv-v0_best.pt holds the same weights as _latest.pt (checked array by array): not stored twice
    for tag, fname in (("w3d_latest", "mocca_envs:Walker3DStepperEnv-v0_latest.pt"), ("w3d_base", "mocca_envs:Walker3DStepperEnv-v0_base.pt"),
                       ("mike_latest", "mocca_envs:MikeStepperEnv-v0_latest.pt")):
        arrs = actor_arrays(os.path.join(MODELS, fname))
        print(tag, str(arrs["actor_class"]), {k: v.shape for k, v in arrs.items() if k.endswith("_w")}, "logstd", arrs["logstd"][:3])
        if "--dry" in sys.argv:
            continue
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "policies", "%s.npz" % tag), source=np.array(fname), **arrs)


if __name__ == "__main__":
    main()
