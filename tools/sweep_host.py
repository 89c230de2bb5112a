"""Sweep the host-buffer path's knobs (diagnostic): SPL_HOST_CHUNKS x SPL_HOST_THREADS.
usage: python tools/sweep_host.py ENVS 'CHUNKS,THREADS' ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from splendor_gym_b200 import SplendorVecEnv

N = int(sys.argv[1])
for spec in sys.argv[2:]:
    chunks, threads = [int(x) for x in spec.split(",")]
    os.environ["SPL_HOST_CHUNKS"] = str(chunks)
    env = SplendorVecEnv(N, device="cuda:0", seed=1, shuffle="philox", autoreset=True)
    env.lib.spl_host_set_threads(threads)
    for dt in (torch.int32, torch.uint8):
        _, info = env.reset_host(obs_dtype=dt, sample_next=True)
        act = env._host["next_action"].numpy().copy()
        for _ in range(10):
            _, _, _, _, info = env.step_host(act, obs_dtype=dt, sample_next=True)
            np.copyto(act, info["next_action"].numpy())
        t0 = time.perf_counter()
        reps = 100
        for _ in range(reps):
            _, _, _, _, info = env.step_host(act, obs_dtype=dt, sample_next=True)
            np.copyto(act, info["next_action"].numpy())
        el = time.perf_counter() - t0
        print(f"envs={N} chunks={chunks:2d} threads={threads:2d} obs={str(dt):12s} {1e6 * el / reps:7.1f} us per lock-step  {N * reps / el / 1e6:7.2f} M env-steps/s", flush=True)
    env.close()
