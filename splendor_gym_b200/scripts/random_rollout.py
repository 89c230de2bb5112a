"""`python -m splendor_gym_b200.scripts.random_rollout --episodes 3` -- the reference's
splendor_gym/scripts/random_rollout.py:13-30 on the CUDA engine (BASELINE config 1), plus `--envs N` to play
N games at once with the batched API."""
from __future__ import annotations

import argparse

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--envs", type=int, default=0, help="> 0: play this many games in lock-step on the device instead")
    args = ap.parse_args(argv)

    if args.envs > 0:
        import torch

        from ..vec_env import SplendorVecEnv

        env = SplendorVecEnv(args.envs, seed=args.seed, shuffle="philox", autoreset=True)
        env.reset()
        actions = env.sample_random_actions()
        steps = 0
        while int(env.stats[0]) < args.episodes * args.envs and steps < 500 * args.episodes:
            env.step(actions, sample_next=True)
            actions = env.next_action
            steps += 1
        st = dict(zip(("episodes", "p0_wins", "p1_wins", "tie_draws", "limit_draws", "nolegal_draws", "sum_moves", "sum_winner_prestige"),
                      env.stats.cpu().tolist()))
        print(f"{args.envs} envs x {steps} lock-steps: {st}")
        return st

    from ..envs import SplendorEnv

    def play_one(env, seed: int, cap: int = 500):
        """One game of uniformly random legal moves on the facade; returns (moves, last reward)."""
        _, info = env.reset(seed=seed)
        last, moves = 0.0, 0
        for moves in range(1, cap + 1):
            legal = np.flatnonzero(info["action_mask"])
            if legal.size == 0:
                return moves - 1, last
            _, last, terminated, truncated, info = env.step(int(np.random.choice(legal)))
            if terminated or truncated:
                break
        return moves, last

    env = SplendorEnv(num_players=2)
    outcomes = [play_one(env, args.seed + ep) for ep in range(args.episodes)]
    for ep, (steps, reward) in enumerate(outcomes):
        print(f"Episode {ep}: steps={steps} reward={reward}")
    wins = sum(1 for _, r in outcomes if r > 0)
    print(f"Wins: {wins}/{args.episodes}")
    return wins


if __name__ == "__main__":
    main()
