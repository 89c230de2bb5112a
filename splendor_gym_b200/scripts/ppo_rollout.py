"""Rollout collection of ppo_splendor.py:202-297 with the MLP policy in the loop (BASELINE config 3), with the
environment on the device: observations / masks never leave HBM, the per-env Python loop and both host<->device
copies of the reference (:221-225, :235-250) are gone.  The policy network is the reference's architecture
(ppo_splendor.py:41-59: 297 -> 256 -> 256 -> {45, 1}, tanh) in plain PyTorch -- it is the caller of the hot path,
not part of it.  The agent samples from the masked categorical (:27-38), the opponent plays the masked argmax
of the same network (scripts/eval_suite.py:131-141), turns are DualStepNativeWrapper.dual_step (:90-193).
"""
from __future__ import annotations

import argparse
import json
import time

import torch
import torch.nn as nn

from ..vec_env import SplendorVecEnv


class ActorCritic(nn.Module):
    def __init__(self, obs_dim: int = 297, act_dim: int = 45):
        super().__init__()
        self.critic = nn.Sequential(nn.Linear(obs_dim, 256), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(), nn.Linear(256, 1))
        self.actor = nn.Sequential(nn.Linear(obs_dim, 256), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(), nn.Linear(256, act_dim))


def pad_first_layer(seq: nn.Sequential, pitch: int = 304) -> nn.Sequential:
    """Same network on the env's fp16 policy input ``[N, pitch]`` (obs_format='f16'): the first Linear gets ``pitch - 297`` zero
    input columns, so its output is unchanged while K becomes a multiple of 8 (tensor-core friendly, 16-byte rows)."""
    first = seq[0]
    padded = nn.Linear(pitch, first.out_features, device=first.weight.device, dtype=first.weight.dtype)
    with torch.no_grad():
        padded.weight.zero_()
        padded.weight[:, : first.in_features].copy_(first.weight)
        padded.bias.copy_(first.bias)
    return nn.Sequential(padded, *list(seq)[1:])


def pad_head(seq: nn.Sequential, out_features: int = 48) -> nn.Sequential:
    """Output layer padded with zero rows to a multiple of 8 (slice ``[:, :45]`` afterwards)."""
    last = seq[-1]
    padded = nn.Linear(last.in_features, out_features, device=last.weight.device, dtype=last.weight.dtype)
    with torch.no_grad():
        padded.weight.zero_()
        padded.bias.zero_()
        padded.weight[: last.out_features].copy_(last.weight)
        padded.bias[: last.out_features].copy_(last.bias)
    return nn.Sequential(*list(seq)[:-1], padded)


def policy_input(env: SplendorVecEnv, dtype):
    """What the MLP reads: the env's own fp16 tensor (cast fused into the step kernel) or the reference's cast (ppo_splendor.py:221)."""
    return env.obs_f16 if env.obs_format == "f16" else env.obs.to(dtype)


def advantages(env: SplendorVecEnv, net: "ActorCritic", buffers, gamma: float = 0.99, gae_lambda: float = 0.95, dtype=torch.float32):
    """GAE over the collected segment on the device (ppo_splendor.py:299-314)."""
    from ..policy import gae

    with torch.no_grad():
        last_values = net.critic(policy_input(env, dtype)).float().squeeze(1)
    return gae(buffers["rewards"], buffers["values"], buffers["terminals"], last_values, gamma, gae_lambda)


@torch.no_grad()
def collect(env: SplendorVecEnv, net: ActorCritic, num_steps: int, buffers=None, dtype=torch.float32):
    """num_steps dual-steps for every env; fills (and returns) rollout buffers shaped like ppo_splendor.py:210-217."""
    n, dev = env.n, env.device
    if buffers is None:
        buffers = dict(
            obs=torch.zeros((num_steps, n, 297), dtype=env.obs.dtype, device=dev), masks=torch.zeros((num_steps, n, 45), dtype=torch.int8, device=dev),
            actions=torch.zeros((num_steps, n), dtype=torch.int32, device=dev), logprobs=torch.zeros((num_steps, n), device=dev),
            values=torch.zeros((num_steps, n), device=dev), rewards=torch.zeros((num_steps, n), device=dev),
            terminals=torch.zeros((num_steps, n), dtype=torch.bool, device=dev),
        )

    from ..policy import masked_sample

    def opponent(obs, mask):  # model_greedy_policy_from: argmax of the masked logits (scripts/eval_suite.py:131-141)
        return masked_sample(net.actor(policy_input(env, dtype))[:, :45], mask, greedy=True, want_logprob=False)[0]

    for t in range(num_steps):
        x = policy_input(env, dtype)
        # masked categorical sample + log-prob in one kernel (ppo_splendor.py:27-38,54-59)
        action, logprob, _ = masked_sample(net.actor(x)[:, :45], env.mask, t=env._t)
        buffers["obs"][t].copy_(env.obs)
        buffers["masks"][t].copy_(env.mask)
        buffers["actions"][t].copy_(action)
        buffers["logprobs"][t].copy_(logprob)
        buffers["values"][t].copy_(net.critic(x).float().squeeze(1))
        _, agent_r, _, _, done, _ = env.dual_step(action, opponent)
        buffers["rewards"][t].copy_(agent_r)
        buffers["terminals"][t].copy_(done)
    return buffers


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=262144)
    ap.add_argument("--num-steps", type=int, default=128)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--checkpoint", default=None, help="state_dict of the reference's ActorCritic (runs/ppo_splendor/ppo_splendor_latest.pt)")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--obs-format", default="int32", choices=["int32", "f16"],
                    help="f16: the step kernel emits the fp16 [N,304] policy input itself (policy runs in fp16, padded layers)")
    ap.add_argument("--shuffle", default="philox", choices=["philox", "mt19937"],
                    help="mt19937: every deal bit-identical to the reference's initial_state(seed) (prefetched deals)")
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args(argv)
    torch.manual_seed(args.seed)
    dev = torch.device("cuda")
    net = ActorCritic().to(dev)
    if args.checkpoint:
        net.load_state_dict(torch.load(args.checkpoint, map_location=dev))
    dtype = torch.float16 if args.obs_format == "f16" else (torch.bfloat16 if args.bf16 else torch.float32)
    net = net.to(dtype).eval()
    if args.obs_format == "f16":
        net.actor, net.critic = pad_head(pad_first_layer(net.actor)), pad_first_layer(net.critic)
    env = SplendorVecEnv(args.num_envs, seed=args.seed, shuffle=args.shuffle, autoreset=True, obs_format=args.obs_format)
    env.reset()
    # rollout buffers in chunks so that 262,144 envs x 128 steps (40 GB of int32 observations) is not required at once
    chunk = max(1, min(args.num_steps, int(8e9 // (args.num_envs * 297 * env.obs.element_size()))))
    buf = collect(env, net, chunk, dtype=dtype)
    torch.cuda.synchronize()
    env_ms = 0.0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 0
    for _ in range(args.repeats):
        done = 0
        while done < args.num_steps:
            k = min(chunk, args.num_steps - done)
            collect(env, net, k, buffers=buf, dtype=dtype)
            if k == chunk:
                advantages(env, net, buf, dtype=dtype)
            done += k
            steps += k
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    agent_steps = steps * args.num_envs
    out = {"config": "PPO-MLP self-play rollout (BASELINE configs[2])", "num_envs": args.num_envs, "dual_steps": steps,
           "agent_steps_per_s": agent_steps / (ms * 1e-3), "env_steps_per_s_upper": 2 * agent_steps / (ms * 1e-3),
           "ms_per_dual_step": ms / steps, "policy_dtype": str(dtype), "obs_format": args.obs_format, "wall_s": time.perf_counter() - t0,
           "episode_stats": env.stats.cpu().tolist()}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
