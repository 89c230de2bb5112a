import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from splendor_gym_b200 import SplendorVecEnv
N = int(sys.argv[1]); T = 16
dev = torch.device("cuda", 0)
env = SplendorVecEnv(N, device=dev, seed=1, shuffle="philox", autoreset=True)
obs = torch.zeros((T, N, 297), dtype=torch.int32, device=dev); mask = torch.zeros((T, N, 45), dtype=torch.int8, device=dev)
rew = torch.zeros((T, N), dtype=torch.float32, device=dev); term = torch.zeros((T, N), dtype=torch.uint8, device=dev)
act = torch.zeros((T + 1, N), dtype=torch.int32, device=dev)
env.t_base = torch.zeros(1, dtype=torch.int64, device=dev)
env.reset(); env.sample_random_actions(out=act[0])
def seg():
    for t in range(T):
        env._t = t
        env.step(act[t], out_obs=obs[t], out_mask=mask[t], out_reward=rew[t], out_terminated=term[t], out_next_action=act[t + 1])
    act[0].copy_(act[T]); env.t_base += T
for spec in sys.argv[2:]:
    pers, wpc, order = (spec.split(",") + ["1", "0"])[:3]
    os.environ["SPL_STEP_PERSISTENT"] = pers
    os.environ["SPL_STEP_WPC"] = wpc
    os.environ["SPL_STEP_ORDER"] = order
    for _ in range(3): seg()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); seg(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); med = ts[5]
    print(f"envs={N} persistent={pers} wpc={wpc} order={order}: {1e3*med/T:.1f} us per lock-step, {N*T/med/1e6:.3f} G env-steps/s, {N*T*1375/med/1e6:.0f} GB/s", flush=True)
