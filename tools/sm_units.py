"""Diagnostic (needs a -DSPL_DEBUG_SM_UNITS build selected with SPL_LIB): how many warp-lock-steps each SM processed
in a rollout launch -- shows whether the work queue shifts work away from the SMs that store slowly."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from splendor_gym_b200 import SplendorVecEnv

N, T = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
env = SplendorVecEnv(N, device=dev, seed=1, shuffle="philox", autoreset=True)
obs = torch.zeros((T, N, 297), dtype=torch.int32, device=dev)
mask = torch.zeros((T, N, 45), dtype=torch.int8, device=dev)
rew = torch.zeros((T, N), dtype=torch.float32, device=dev)
term = torch.zeros((T, N), dtype=torch.uint8, device=dev)
act = torch.zeros((T + 1, N), dtype=torch.int32, device=dev)
env.reset()
env.sample_random_actions(out=act[0])
out = np.zeros(256, np.uint32)
for rep in range(3):
    env.rollout_random(T, act[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=act)
    act[0].copy_(act[T])
    torch.cuda.synchronize()
    env.lib.spl_debug_sm_units(out.ctypes.data_as(C.c_void_p), 1)
u = out[:148].astype(np.int64)
print(f"envs={N} T={T}: warp-lock-steps per SM: min {u.min()} median {int(np.median(u))} max {u.max()} mean {u.mean():.0f}")
print(" ".join(str(x) for x in u))
