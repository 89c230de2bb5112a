from .dual_step_native import DualStepNativeWrapper
from .selfplay import SelfPlayWrapper, random_opponent, vec_selfplay_step

__all__ = ["SelfPlayWrapper", "DualStepNativeWrapper", "random_opponent", "vec_selfplay_step"]
