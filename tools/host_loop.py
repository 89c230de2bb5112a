"""A short host-buffer loop (SplendorVecEnv.step_host) for profiling: python tools/host_loop.py [ENVS] [STEPS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from splendor_gym_b200 import SplendorVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
env = SplendorVecEnv(n, device="cuda:0", seed=1, shuffle="philox", autoreset=True)
_, info = env.reset_host(sample_next=True)
act = env._host["next_action"].numpy().copy()
for _ in range(steps):
    _, _, _, _, info = env.step_host(act, sample_next=True)
    np.copyto(act, info["next_action"].numpy())
print("host loop done:", env.host_stats())
env.close()
