"""Sweep the rollout kernel's launch knobs (diagnostic): SPL_WPC, SPL_ROLLOUT_CHUNK, SPL_ROLLOUT_CTAS_PER_SM, SPL_ROLLOUT_SYNC.
usage: python tools/sweep_rollout.py ENVS STEPS 'WPC,CHUNK,CTAS_PER_SM,SYNC' ...   (-1 = library default)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from splendor_gym_b200 import SplendorVecEnv

N, T = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
env = SplendorVecEnv(N, device=dev, seed=20261018, shuffle="philox", autoreset=True)
obs = torch.zeros((T, N, 297), dtype=torch.int32, device=dev)
mask = torch.zeros((T, N, 45), dtype=torch.int8, device=dev)
rew = torch.zeros((T, N), dtype=torch.float32, device=dev)
term = torch.zeros((T, N), dtype=torch.uint8, device=dev)
act = torch.zeros((T + 1, N), dtype=torch.int32, device=dev)
env.t_base = torch.zeros(1, dtype=torch.int64, device=dev)
env.reset()
env.sample_random_actions(out=act[0])
bytes_per_launch = N * T * (1188 + 45 + 4 + 1 + 4) + N * 132
names = ["SPL_WPC", "SPL_ROLLOUT_CHUNK", "SPL_ROLLOUT_CTAS_PER_SM", "SPL_ROLLOUT_SYNC"]


def seg():
    env._t = 0
    env.rollout_random(T, act[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=act)
    act[0].copy_(act[T])
    env.t_base += T


for spec in sys.argv[3:]:
    vals = [int(x) for x in spec.split(",")]
    for k, v in zip(names, vals):
        if v < 0:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    for _ in range(3):
        seg()
    torch.cuda.synchronize()
    times = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        seg()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    med = times[len(times) // 2]
    print(f"envs={N} T={T} wpc,chunk,ctas/sm,sync={spec:>14s}  median {med:.3f} ms  best {times[0]:.3f} ms  "
          f"{N * T / med / 1e6:.3f} G env-steps/s  {bytes_per_launch / med / 1e6:.0f} GB/s", flush=True)
