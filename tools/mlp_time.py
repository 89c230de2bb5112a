"""Diagnostic: where does the PPO-MLP forward (ppo_splendor.py:41-59) spend its time at 262,144 envs?"""
import torch, torch.nn as nn, sys
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = "cuda"
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
obs = torch.randint(0, 8, (N, 297), dtype=torch.int32, device=dev)
for dt in (torch.float32, torch.bfloat16):
    for K in (297, 304, 320):
        actor = nn.Sequential(nn.Linear(K, 256), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(), nn.Linear(256, 45)).to(dev).to(dt)
        x = torch.zeros((N, K), dtype=dt, device=dev)
        with torch.no_grad():
            cast = t(lambda: obs.to(dt))
            fwd = t(lambda: actor(x))
            l1 = t(lambda: actor[0](x))
            h = actor[1](actor[0](x))
            tanh = t(lambda: actor[1](h))
            l2 = t(lambda: actor[2](h))
            l3 = t(lambda: actor[4](h))
        print(f"N={N} {str(dt):15s} K={K}: cast {cast:.3f} ms  actor fwd {fwd:.3f} ms  [L1 {l1:.3f}  tanh {tanh:.3f}  L2 {l2:.3f}  L3 {l3:.3f}]", flush=True)
