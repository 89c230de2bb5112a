"""Diagnostic: kernel-time breakdown of one PPO-MLP dual-step (scripts/ppo_rollout.py, obs_format=f16) via torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from splendor_gym_b200 import SplendorVecEnv
from splendor_gym_b200.scripts.ppo_rollout import ActorCritic, collect, pad_first_layer, pad_head

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
net = ActorCritic().cuda().half().eval()
net.actor, net.critic = pad_head(pad_first_layer(net.actor)), pad_first_layer(net.critic)
env = SplendorVecEnv(N, seed=1, shuffle="philox", autoreset=True, obs_format="f16")
env.reset()
buf = collect(env, net, 8, dtype=torch.float16)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    collect(env, net, 8, buffers=buf, dtype=torch.float16)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 8e3:.3f} ms per dual-step")
for e in rows[:22]:
    print(f"{e.device_time_total / 8e3:8.3f} ms/dual-step  x{e.count / 8:5.1f}  {e.key[:110]}")
