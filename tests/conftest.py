import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import binding
    binding.build()
    return binding


def make_oracle(kind, n, seed=8, threads=0, params=None, terrain_mode=None):
    """CPU oracle with the same model / params / RNG streams as BatchedStepperEnv(seed, n)."""
    from oracle.binding import Oracle
    from steppingstone_b200 import model, seeding
    M = model.make_model(kind)
    P = model.make_params(kind, params, terrain_mode)
    return Oracle(M, P, n, seeding.env_streams(seed, 0, n), seeding.aux_streams(kind, seed, 0, n), threads=threads)


def make_gpu_env(kind, n, seed=8, output="numpy", params=None, terrain_mode=None):
    from steppingstone_b200 import BatchedStepperEnv
    env_id = "mocca_envs:Walker3DStepperEnv-v0" if kind == 0 else "mocca_envs:MikeStepperEnv-v0"
    return BatchedStepperEnv(env_id, seed, n, output=output, params=params, terrain_mode=terrain_mode)


def random_actions(rng, n, sd=0.22):
    """clip(N(0, e^-1.5), -1, 1): the initial policy's action distribution (controller.py:21)."""
    return np.clip(rng.randn(n, 21) * sd, -1, 1).astype(np.float32)
