"""Masked PPO self-play on the device-resident Splendor engine -- the training loop of the reference's
ppo_splendor.py (:62-412) with the rollout on the GPU.

What is ours here is the rollout side: batched `dual_step`, masked
sampling / log-probs (`spl_masked_sample`), GAE (`spl_gae`), scripted or pooled opponents with per-env opponent ids
and grouped batched inference (ppo_splendor.py:135-143,366-370), batched evaluation.  The PPO update itself
(:327-361) is the reference's algorithm in plain PyTorch -- it is the consumer of the hot path, not part of it --
kept so that the package trains end to end with the reference's hyper-parameters.

    python -m splendor_gym_b200.scripts.ppo_train --num-envs 4096 --total-timesteps 4000000
"""
from __future__ import annotations

import argparse
import copy
import json
import time

import torch
import torch.nn as nn

from ..policy import bot_policy, gae, masked_sample
from ..vec_env import SplendorVecEnv
from .eval_suite import eval_vs_opponent
from .ppo_rollout import ActorCritic


class OpponentPool:
    """Per-env opponents with per-episode resampling, the batched form of ppo_splendor.py:135-143,366-370:

    * every env holds one opponent for the whole episode; when its episode ends a new one is drawn -- the CURRENT policy
      with probability ``p_current`` (always, while the pool is empty), else a uniformly random frozen snapshot (:137-143);
    * ``add_snapshot`` appends a frozen copy of the actor and drops the oldest beyond ``pool_size`` (:366-370).  An env
      that is still playing against a dropped snapshot keeps it until its episode ends, as the reference's closure does
      (training_utils.py:263-276 builds the frozen policy once, in ``wrapper.reset()``);
    * opponents act greedily: argmax of the masked logits (scripts/eval_suite.py:131-141, training_utils.py:270-276).

    Opponent ids: -1 = the current policy, k >= 0 = the k-th snapshot ever taken.  ``act`` runs one batched forward per
    distinct id present (grouped inference)."""

    def __init__(self, net: ActorCritic, n: int, device, p_current: float = 0.25, pool_size: int = 12):
        self.net, self.n, self.device = net, n, device
        self.p_current, self.pool_size = p_current, pool_size
        self.pool: list[int] = []             # ids of the snapshots new episodes can draw, oldest first
        self.models: dict[int, nn.Module] = {}  # id -> frozen actor (also dropped ones that an env still plays against)
        self.taken = 0
        self.opp_id = torch.full((n,), -1, dtype=torch.int64, device=device)

    @property
    def snapshots(self) -> list:
        return [self.models[k] for k in self.pool]

    def add_snapshot(self):
        snap = copy.deepcopy(self.net.actor).eval()
        for p in snap.parameters():
            p.requires_grad_(False)
        self.models[self.taken] = snap
        self.pool.append(self.taken)
        self.taken += 1
        if len(self.pool) > self.pool_size:
            self.pool.pop(0)
        in_use = set(self.opp_id.unique().tolist())
        for k in [k for k in self.models if k not in self.pool and k not in in_use]:
            del self.models[k]

    def resample(self, done: torch.Tensor):
        """New opponent for every env whose episode just ended (the reference resamples in wrapper.reset())."""
        if not self.pool:
            return
        cur = torch.rand(self.n, device=self.device) < self.p_current
        ids = torch.tensor(self.pool, dtype=torch.int64, device=self.device)
        pick = ids[torch.randint(0, len(self.pool), (self.n,), device=self.device)]
        new = torch.where(cur, torch.full_like(pick, -1), pick)
        self.opp_id = torch.where(done, new, self.opp_id)

    @torch.no_grad()
    def act(self, obs: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        x = obs.float()
        logits = torch.empty((self.n, 45), dtype=torch.float32, device=self.device)
        for k in self.opp_id.unique().tolist():
            sel = (self.opp_id == k).nonzero(as_tuple=True)[0]
            model = self.net.actor if k < 0 else self.models[k]
            logits[sel] = model(x[sel]).float()
        return masked_sample(logits, mask, greedy=True, want_logprob=False)[0]  # model_greedy_policy_from


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-timesteps", type=int, default=2_000_000)
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--num-steps", type=int, default=128)
    ap.add_argument("--gamma", type=float, default=0.999)
    ap.add_argument("--gae-lambda", type=float, default=0.95)
    ap.add_argument("--lr", type=float, default=2.5e-4)
    ap.add_argument("--ent-coef", type=float, default=0.03)
    ap.add_argument("--ent-coef-final", type=float, default=0.01)
    ap.add_argument("--vf-coef", type=float, default=0.5)
    ap.add_argument("--clip-coef", type=float, default=0.2)
    ap.add_argument("--vclip", type=float, default=0.2)
    ap.add_argument("--update-epochs", type=int, default=4)
    ap.add_argument("--minibatch-size", type=int, default=16384)
    ap.add_argument("--target-kl", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--opponent", default="pool", choices=["pool", "random", "greedy_v1", "basic", "greedy_v2"])
    ap.add_argument("--pool-size", type=int, default=12)
    ap.add_argument("--snapshot-every-updates", type=int, default=10)
    ap.add_argument("--p-current", type=float, default=0.25)
    ap.add_argument("--eval-every-updates", type=int, default=10)
    ap.add_argument("--eval-games", type=int, default=2000)
    ap.add_argument("--save-path", default=None)
    args = ap.parse_args(argv)

    torch.manual_seed(args.seed)
    dev = torch.device("cuda")
    n, T = args.num_envs, args.num_steps
    net = ActorCritic().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr, eps=1e-5)
    env = SplendorVecEnv(n, seed=args.seed, shuffle="philox", autoreset=True)
    env.reset()
    pool = OpponentPool(net, n, dev, args.p_current, args.pool_size) if args.opponent == "pool" else None
    opponent = pool.act if pool is not None else bot_policy(args.opponent)

    obs_b = torch.zeros((T, n, 297), dtype=torch.int32, device=dev)
    mask_b = torch.zeros((T, n, 45), dtype=torch.int8, device=dev)
    act_b = torch.zeros((T, n), dtype=torch.int32, device=dev)
    logp_b = torch.zeros((T, n), device=dev)
    val_b = torch.zeros((T, n), device=dev)
    rew_b = torch.zeros((T, n), device=dev)
    done_b = torch.zeros((T, n), dtype=torch.bool, device=dev)

    num_updates = max(1, args.total_timesteps // (n * T))
    log, t_start, global_step = [], time.perf_counter(), 0
    for update in range(num_updates):
        t0 = time.perf_counter()
        with torch.no_grad():
            for t in range(T):  # rollout collection, ppo_splendor.py:219-297, nothing leaves the device
                x = env.obs.float()
                action, logp, _ = masked_sample(net.actor(x), env.mask, t=global_step + t, key=args.seed)
                obs_b[t].copy_(env.obs)
                mask_b[t].copy_(env.mask)
                act_b[t].copy_(action)
                logp_b[t].copy_(logp)
                val_b[t].copy_(net.critic(x).squeeze(1))
                _, agent_r, _, _, done, _ = env.dual_step(action, opponent)
                rew_b[t].copy_(agent_r)
                done_b[t].copy_(done)
                if pool is not None:
                    pool.resample(done)
            last_v = net.critic(env.obs.float()).squeeze(1)
            adv, ret = gae(rew_b, val_b, done_b, last_v, args.gamma, args.gae_lambda)
        torch.cuda.synchronize()
        t_roll = time.perf_counter() - t0
        global_step += n * T

        # ---- PPO update (ppo_splendor.py:327-361), plain PyTorch
        B = n * T
        b_obs, b_mask = obs_b.view(B, 297), mask_b.view(B, 45)
        b_act, b_logp, b_val = act_b.view(B).long(), logp_b.view(B), val_b.view(B)
        b_ret, b_adv = ret.view(B), adv.view(B)
        b_adv = (b_adv - b_adv.mean()) / (b_adv.std() + 1e-8)
        ent_now = args.ent_coef + (args.ent_coef_final - args.ent_coef) * (update / max(1, num_updates - 1))
        mb = min(args.minibatch_size, B)
        stop = False
        for _ in range(args.update_epochs):
            perm = torch.randperm(B, device=dev)
            for s in range(0, B, mb):
                idx = perm[s:s + mb]
                x = b_obs[idx].float()
                logits = net.actor(x)
                illegal = b_mask[idx] < 1
                any_legal = (~illegal).any(dim=1, keepdim=True)
                logits = torch.where(illegal & any_legal, torch.full_like(logits, float("-inf")), logits)
                dist = torch.distributions.Categorical(logits=logits)
                new_logp = dist.log_prob(b_act[idx])
                ratio = (new_logp - b_logp[idx]).exp()
                a = b_adv[idx]
                pol_loss = -torch.min(ratio * a, torch.clamp(ratio, 1 - args.clip_coef, 1 + args.clip_coef) * a).mean()
                v = net.critic(x).squeeze(1)
                v_clip = b_val[idx] + torch.clamp(v - b_val[idx], -args.vclip, args.vclip)
                v_loss = 0.5 * torch.max((v - b_ret[idx]).pow(2), (v_clip - b_ret[idx]).pow(2)).mean()
                loss = pol_loss + args.vf_coef * v_loss - ent_now * dist.entropy().mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(net.parameters(), 0.5)
                opt.step()
                if args.target_kl > 0 and float((b_logp[idx] - new_logp.detach()).mean()) > args.target_kl:
                    stop = True
                    break
            if stop:
                break
        if pool is not None and (update + 1) % max(1, args.snapshot_every_updates) == 0:
            pool.add_snapshot()
        torch.cuda.synchronize()
        rec = {"update": update + 1, "global_step": global_step, "rollout_agent_steps_per_s": n * T / t_roll,
               "update_s": time.perf_counter() - t0 - t_roll, "episodes": int(env.stats[0]), "loss": float(loss.detach())}
        if (update + 1) % max(1, args.eval_every_updates) == 0 or update + 1 == num_updates:
            @torch.no_grad()
            def greedy(obs, mask):
                return masked_sample(net.actor(obs.float()), mask, greedy=True, want_logprob=False)[0]
            for opp in ("random", "greedy_v1", "basic"):
                rec[f"win_rate_vs_{opp}"] = eval_vs_opponent(greedy, opp, n_games=args.eval_games, seed=update)["win_rate"]
        log.append(rec)
        print(json.dumps(rec))
    if args.save_path:
        torch.save(net.state_dict(), args.save_path)
    print(json.dumps({"total_s": time.perf_counter() - t_start, "timesteps": global_step, "sps": global_step / (time.perf_counter() - t_start)}))
    return log


if __name__ == "__main__":
    main()
