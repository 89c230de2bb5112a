"""Sweep the host-buffer path's knobs (diagnostic): worker threads x push-kernel CTAs x GPU-written share.
usage: python tools/sweep_host.py ENVS 'THREADS,CTAS,DIRECT[,CTA_THREADS[,NIBBLES[,RING_SLOTS]]]' ...     (DIRECT: share in [0,1], or -1 = feedback)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from splendor_gym_b200 import SplendorVecEnv

N = int(sys.argv[1])
dtypes = (torch.int32, torch.uint8) if os.environ.get("SWEEP_U8") else (torch.int32,)
for spec in sys.argv[2:]:
    parts = spec.split(",") + ["64", "1", "16"][len(spec.split(",")) - 3:]
    threads, ctas, direct, cta_threads, nibbles, ring = parts[:6]
    os.environ["SPL_RING_SLOTS"] = ring
    os.environ["SPL_PUSH_CTAS"] = ctas
    os.environ["SPL_PUSH_THREADS"] = cta_threads
    os.environ["SPL_HOST_NIBBLES"] = nibbles
    if float(direct) >= 0:
        os.environ["SPL_HOST_DIRECT"] = direct
    else:
        os.environ.pop("SPL_HOST_DIRECT", None)
    env = SplendorVecEnv(N, device="cuda:0", seed=1, shuffle="philox", autoreset=True)
    env.lib.spl_host_set_threads(int(threads))
    for dt in dtypes:
        _, info = env.reset_host(obs_dtype=dt, sample_next=True)
        act = env._host["next_action"].numpy().copy()
        for _ in range(40):
            _, _, _, _, info = env.step_host(act, obs_dtype=dt, sample_next=True)
            np.copyto(act, info["next_action"].numpy())
        t0 = time.perf_counter()
        reps = 100
        acc = {}
        for _ in range(reps):
            _, _, _, _, info = env.step_host(act, obs_dtype=dt, sample_next=True)
            np.copyto(act, info["next_action"].numpy())
            for k, v in env.host_stats().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        el = time.perf_counter() - t0
        print(f"envs={N} threads={threads:>2s} ctas={ctas:>3s}x{cta_threads:>3s} nib={nibbles} ring={ring:>3s} direct={direct:>5s} obs={str(dt):12s} {1e6 * el / reps:7.1f} us per lock-step  "
              f"{N * reps / el / 1e6:7.2f} M env-steps/s | call {acc['call_us']:.0f} enq {acc['enqueued_us']:.0f} first {acc['first_group_us']:.0f} "
              f"workers {acc['workers_done_us']:.0f} gpu {acc['gpu_share_done_us']:.0f} share {acc['gpu_written_share']:.3f}", flush=True)
    env.close()
