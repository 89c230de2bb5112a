from .dual_step_native import DualStepNativeWrapper
from .dual_step_selfplay import DualStepSelfPlayWrapper
from .selfplay import SelfPlayWrapper, random_opponent, vec_selfplay_step

__all__ = ["SelfPlayWrapper", "DualStepNativeWrapper", "DualStepSelfPlayWrapper", "random_opponent", "vec_selfplay_step"]
