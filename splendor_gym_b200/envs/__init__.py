"""Single-environment facade (SplendorEnv) over the CUDA engine; the batched environment is splendor_gym_b200.SplendorVecEnv."""
from .splendor_env import SplendorEnv, make

__all__ = ["SplendorEnv", "make"]
