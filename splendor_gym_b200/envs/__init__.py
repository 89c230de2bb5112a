"""Single-environment facade (SplendorEnv) over the CUDA engine; the batched environment is splendor_gym_b200.SplendorVecEnv."""
from . import splendor_env as _impl

SplendorEnv = _impl.SplendorEnv
make = _impl.make

__all__ = ("SplendorEnv", "make")
