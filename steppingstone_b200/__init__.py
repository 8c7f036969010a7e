"""B200-native batched stepping-stone environment (drop-in for the reference's vec-env hot path)."""
from .vec_env import BatchedStepperEnv, make_vec_envs, StepInfo  # noqa: F401

__all__ = ["BatchedStepperEnv", "make_vec_envs", "StepInfo"]
