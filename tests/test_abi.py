"""CPU checks of the C-ABI boundary: the library builds, loads and exports every symbol include/ss_env.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ss_env.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from steppingstone_b200 import native
    lib = ctypes.CDLL(native.build_native())
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libss_b200.so does not export %s" % n
    assert set(names) == set(native.EXPORTS), set(names) ^ set(native.EXPORTS)


def test_struct_layouts_match_header():
    """ctypes mirrors of ss_model / ss_params have the sizes gcc gives the C structs."""
    import subprocess, tempfile
    from steppingstone_b200 import model
    code = '#include <stdio.h>\n#include "ss_env.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(ss_model), sizeof(ss_params), sizeof(ss_state_view));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    from steppingstone_b200 import native
    assert int(out[0]) == ctypes.sizeof(model.SSModel)
    assert int(out[1]) == ctypes.sizeof(model.SSParams)
    assert int(out[2]) == ctypes.sizeof(native.StateView)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from steppingstone_b200 import make_vec_envs, native
    with pytest.raises(native.NativeError):
        make_vec_envs("mocca_envs:Walker3DStepperEnv-v0", 8, 4)


def test_product_never_imports_oracle():
    """The shipped package must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "steppingstone_b200")
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle)|libss_oracle|\bsso_[a-z_]+\s*\(|ss_oracle\.h", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s references the oracle" % f


def _build_c_host(tmp):
    import subprocess, sys
    from steppingstone_b200 import native
    native.build_native()
    exe = os.path.join(tmp, "c_host")
    libdir = os.path.join(ROOT, "steppingstone_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "examples", "c_host.c"), "-L", libdir, "-lss_b200", "-Wl,-rpath," + libdir, "-lm"])
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_tables.py"), os.path.join(tmp, "tables"), "walker3d", "64", "8"],
                          stdout=subprocess.DEVNULL)
    return exe, os.path.join(tmp, "tables")


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_c_host_links_and_fails_loudly_without_gpu(tmp_path):
    """include/ss_env.h is valid C, a plain-C host links against libss_b200.so, and without a GPU it gets an error code
    and a message (no CPU fallback)."""
    import subprocess
    exe, tables = _build_c_host(str(tmp_path))
    res = subprocess.run([exe, tables, "64", "4"], capture_output=True, text=True)
    assert res.returncode == 1 and "no CUDA device" in res.stderr


@pytest.mark.gpu
def test_c_host_runs_on_gpu(tmp_path):
    import subprocess
    exe, tables = _build_c_host(str(tmp_path))
    res = subprocess.run([exe, tables, "64", "40"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "env-steps/s" in res.stdout and "episodes finished" in res.stdout


def test_sample_table_shapes():
    """update_sample_prob accepts what train.py passes: [N,11,11] stacks (and 3-D / 5-D tables), one table shared."""
    import numpy as np
    from steppingstone_b200.vec_env import expand_sample_table
    rng = np.random.RandomState(0)
    t2 = rng.rand(11, 11); t2 /= t2.sum()
    full = expand_sample_table(np.stack([t2] * 7), 7).reshape((11,) * 5)
    assert np.array_equal(full[:, :, 0, 0, 0], t2) and full.sum() == pytest.approx(1.0) and np.count_nonzero(full) == 121
    t3 = rng.rand(11, 11, 11)
    assert np.array_equal(expand_sample_table(t3, 4).reshape((11,) * 5)[:, :, :, 0, 0], t3)
    t5 = rng.rand(11 ** 5)
    assert np.array_equal(expand_sample_table(t5, 3), t5)
    with pytest.raises(NotImplementedError):
        expand_sample_table(np.stack([t2, t2 * 2]), 2)
    with pytest.raises(ValueError):
        expand_sample_table(np.zeros(10), 2)
