"""splendor_gym_b200 -- B200-native batched Splendor engine behind the SplendorEnv API.

Public surface (mirrors splendor_gym/__init__.py:1-6 and splendor_gym/engine/__init__.py:1-13):
    SplendorVecEnv   batched reset/step/action-mask/dual_step on device tensors
    SplendorEnv      single-environment facade with the Gymnasium SplendorEnv call signature
    make             factory (envs/splendor_env.py:129-130)
    TOTAL_ACTIONS, OBSERVATION_DIM
"""
from ._lib import NUM_ACTIONS as TOTAL_ACTIONS
from ._lib import OBS_DIM as OBSERVATION_DIM
from ._lib import SplendorB200Error


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA (build() runs on a CPU box)
    if name == "SplendorVecEnv":
        from .vec_env import SplendorVecEnv

        return SplendorVecEnv
    if name in ("SplendorEnv", "make"):
        from .envs import splendor_env

        return getattr(splendor_env, name)
    raise AttributeError(name)


__all__ = ["SplendorVecEnv", "SplendorEnv", "make", "TOTAL_ACTIONS", "OBSERVATION_DIM", "SplendorB200Error"]
