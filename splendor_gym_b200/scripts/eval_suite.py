"""Batched `eval_vs_opponent` (splendor_gym/scripts/eval_suite.py:162-208): n_games games of an agent policy
(player 0) against a scripted opponent, one environment per game, all stepped in lock-step on the device with
the SelfPlayWrapper reward convention.  Returns the reference's result dict."""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np
import torch

from ..policy import bot_policy
from ..vec_env import SplendorVecEnv
from ..wrappers import vec_selfplay_step


def eval_vs_opponent(agent_policy: Callable, opponent: str = "random", n_games: int = 400, seed: int = 0, device="cuda",
                     shuffle: str = "philox") -> Dict[str, float]:
    """agent_policy(obs, mask) -> int32 actions [n]; opponent: "random" | "greedy_v1" | "basic" | "greedy_v2" or a callable."""
    env = SplendorVecEnv(n_games, device=device, seed=seed, shuffle=shuffle, autoreset=False)
    env.reset()
    opp = bot_policy(opponent) if isinstance(opponent, str) else opponent
    finished = torch.zeros(n_games, dtype=torch.bool, device=env.device)
    result = torch.zeros(n_games, dtype=torch.float32, device=env.device)
    illegal = torch.zeros((), dtype=torch.int64, device=env.device)
    checks = torch.zeros((), dtype=torch.int64, device=env.device)
    for _ in range(200):  # a game has at most 198 moves = 99 agent turns
        live = ~finished
        a = agent_policy(env.obs, env.mask).to(torch.int32)
        chosen = env.mask.gather(1, a.long().clamp(0, 44).view(-1, 1)).view(-1)
        illegal += ((chosen == 0) & live).sum()
        checks += live.sum()
        _, r, done, _, _ = vec_selfplay_step(env, a, opp)
        newly = done & live
        result = torch.where(newly, r, result)
        finished |= done
        if bool(finished.all()):
            break
    rows = env.export_state().cpu().numpy()
    res = result.cpu().numpy()
    wins, losses = int((res > 0).sum()), int((res < 0).sum())
    n = n_games
    p = wins / max(1, n)
    prev = (rows[:, 70] - 1) % 2  # the player who moved last (scripts/eval_suite.py:190-192)
    prestige = np.where(prev == 0, rows[:, 17], rows[:, 40])
    return {"n": n, "wins": wins, "losses": losses, "draws": n - wins - losses, "win_rate": p,
            "win_rate_ci95": float(1.96 * np.sqrt(p * (1 - p) / max(1, n))), "avg_turns": float(rows[:, 71].mean()),
            "avg_prestige": float(prestige.mean()), "illegal_action_rate": float(illegal.item() / max(1, checks.item()))}
