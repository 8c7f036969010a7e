"""CPU: the bench.py contract for the reference arm (the only arm that runs without a GPU) and the b200 arm's loud failure."""
import json
import os
import subprocess
import sys

from conftest import ROOT, has_gpu
import pytest


def _run(args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run(["--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3", "--ref-envs", "64"])
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("Walker3DStepperEnv-v0") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3", "--ref-envs", "64"],
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_b200_arm_fails_loudly_without_a_gpu():
    res = _run(["--gpus", "1", "--steps", "3", "--warmup", "3"])
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)


def test_stdout_carries_only_the_result_line():
    """Libraries write to stdout too (NCCL prints "NCCL version ..." there at NCCL_DEBUG=VERSION/WARN, seen in a
    2-GPU torchrun of this bench): after claim_stdout() anything written to fd 1 -- from Python or from C -- lands on
    stderr and only emit() reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('chatter'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version" in r.stderr and "chatter" in r.stderr
