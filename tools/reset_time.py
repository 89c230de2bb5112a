"""Diagnostic: latency of spl_reset (one launch) by shuffle mode and env count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from splendor_gym_b200 import SplendorVecEnv

for shuffle in ("mt19937", "philox"):
    for n in (1, 32, 64, 1024, 65536):
        env = SplendorVecEnv(n, device="cuda:0", seed=3, shuffle=shuffle)
        for _ in range(3):
            env.reset()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); env.reset(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"{shuffle:8s} n={n:6d}: reset {ts[len(ts)//2]:8.1f} us (min {ts[0]:.1f})", flush=True)
