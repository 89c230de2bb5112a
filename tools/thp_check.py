"""Are the host result arrays of step_host backed by transparent huge pages?  (diagnostic)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from splendor_gym_b200 import SplendorVecEnv


def rollup():
    out = {}
    for line in open("/proc/self/smaps_rollup"):
        k, _, v = line.partition(":")
        if k in ("Rss", "AnonHugePages", "Anonymous"):
            out[k] = v.strip()
    return out


print("before:", rollup())
env = SplendorVecEnv(65536, device="cuda:0", seed=1, shuffle="philox", autoreset=True)
_, info = env.reset_host(sample_next=True)
print("after host arrays:", rollup())
obs = env._host["obs"]
addr = obs.data_ptr()
for blk in open("/proc/self/smaps").read().split("\n\n") if False else []:
    pass
cur = None
for line in open("/proc/self/smaps"):
    parts = line.split()
    if len(parts) >= 5 and "-" in parts[0] and parts[0][0] in "0123456789abcdef":
        lo, hi = (int(x, 16) for x in parts[0].split("-"))
        cur = (lo, hi) if lo <= addr < hi else None
        if cur:
            print("mapping of the obs array:", line.strip(), f"({(hi - lo) >> 20} MB)")
    elif cur and parts and parts[0] in ("AnonHugePages:", "Rss:", "KernelPageSize:", "MMUPageSize:", "THPeligible:"):
        print("   ", line.strip())
