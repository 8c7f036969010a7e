"""Minimal stand-in for gym (<0.22) -- GOLDEN-GENERATION TOOLING ONLY (tools/gen_golden.py)."""
from . import spaces, utils, envs, core  # noqa: F401
from .core import Env, Wrapper  # noqa: F401
from .wrappers import TimeLimit  # noqa: F401


def make(*a, **k):
    raise NotImplementedError
