"""Diagnostic (needs a -DSPL_DEBUG_PHASES build selected with SPL_LIB:
  cd splendor_gym_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --compiler-options -fPIC,-pthread -shared
      -DSPL_DEBUG_PHASES -o ../libsplendor_b200_phases.so spl_kernels.cu spl_policy.cu spl_host.cu spl_host_expand.cpp -lpthread
  SPL_LIB=splendor_gym_b200/libsplendor_b200_phases.so python tools/step_phases.py 65536): per-warp SM-clock stamps at the phase boundaries
of the single-step kernel -> where one lock-step's time goes (table staging, state load, rules, mask, encode, stores)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from splendor_gym_b200 import SplendorVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 24
dev = torch.device("cuda", 0)
env = SplendorVecEnv(N, device=dev, seed=1, shuffle="philox", autoreset=True)
obs = torch.zeros((T, N, 297), dtype=torch.int32, device=dev)
mask = torch.zeros((T, N, 45), dtype=torch.int8, device=dev)
rew = torch.zeros((T, N), dtype=torch.float32, device=dev)
term = torch.zeros((T, N), dtype=torch.uint8, device=dev)
act = torch.zeros((T + 1, N), dtype=torch.int32, device=dev)
env.reset()
env.sample_random_actions(out=act[0])
warps = min(65536, (N + 31) // 32)
out = np.zeros((13, warps), np.uint64)
names = ["entry", "tables staged", "state+action landed", "rules (step+reset) done", "state/reward stored", "next mask done",
         "mask tile + sample done", "obs encoded", "obs tile stores issued", "end"]
acc = []
for t in range(T):
    env._t = t
    env.step(act[t], out_obs=obs[t], out_mask=mask[t], out_reward=rew[t], out_terminated=term[t], out_next_action=act[t + 1])
    torch.cuda.synchronize()
    env.lib.spl_debug_phases(out.ctypes.data_as(C.c_void_p), warps)
    if t >= 8:
        acc.append(out.copy())
for a in acc[-3:]:
    c = a[:10].astype(np.int64)
    g0, g1, sm = a[10].astype(np.int64), a[12].astype(np.int64), a[11].astype(np.int64)
    print(f"--- lock-step: kernel span by globaltimer {(g1.max() - g0.min()) / 1e3:.1f} us; first warp starts at 0, last warp starts at "
          f"{(g0.max() - g0.min()) / 1e3:.1f} us; warps per SM min/max {np.bincount(sm).min()}/{np.bincount(sm).max()}")
    for k in range(1, 10):
        d = c[k] - c[k - 1]
        print(f"  {names[k]:28s} cycles: mean {d.mean():8.0f}  p10 {np.percentile(d, 10):8.0f}  p50 {np.percentile(d, 50):8.0f}  p90 {np.percentile(d, 90):8.0f}  max {d.max():8.0f}")
    tot = c[9] - c[0]
    print(f"  {'warp total':28s} cycles: mean {tot.mean():8.0f}  p10 {np.percentile(tot, 10):8.0f}  p50 {np.percentile(tot, 50):8.0f}  p90 {np.percentile(tot, 90):8.0f}  max {tot.max():8.0f}")
    # per-warp end time relative to the kernel start (globaltimer)
    e = (g1 - g0.min()) / 1e3
    s = (g0 - g0.min()) / 1e3
    # what the slow tail is made of: the last warps to finish, and the span if every warp's rules phase had taken the median
    rules = c[3] - c[2]
    clk = (c[9] - c[0]).astype(np.float64) / np.maximum((g1 - g0).astype(np.float64), 1.0)  # cycles per ns, per warp
    est = (g1 - g0.min()) / 1e3 - np.maximum(rules - np.median(rules), 0) / np.median(clk) / 1e3
    late = np.argsort(e)[-max(1, len(e) // 100):]
    print(f"  slowest 1 % of warps: rules {rules[late].mean():.0f} cycles (all: {rules.mean():.0f}), obs stores {(c[8] - c[7])[late].mean():.0f} (all: {(c[8] - c[7]).mean():.0f}); "
          f"span with every rules phase capped at its median: {est.max():.1f} us")
    print(f"  warp start us: p50 {np.percentile(s, 50):.1f} p90 {np.percentile(s, 90):.1f} max {s.max():.1f} | warp end us: p10 {np.percentile(e, 10):.1f} p50 {np.percentile(e, 50):.1f} p90 {np.percentile(e, 90):.1f} max {e.max():.1f}")
