"""Compact trajectory format: a rollout is fully determined by
(seed_base, env_offset, shuffle mode, episode counters, packed start state) + one int32 action per env-step, i.e.
~4 bytes per step instead of the 1,242 bytes of observation / mask / reward it regenerates.  `record_random` logs a
random-policy rollout, `replay` re-executes a log on the device and returns (or streams) every observation -- the
dataset-export path for offline RL mentioned in the reference's README -- and `describe_action` renders an action
the way the reference's game logger does (scripts/game_logger.py:98-170) for debugging."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch

from .engine.encode import TAKE3_COMBOS
from .vec_env import SplendorVecEnv

_ABBR = "wbgrk"  # white blue green red black, the abbreviations of scripts/game_logger.py:53 (gold is "G")


def _card_text(card) -> str:
    """``g-1pt-2b3r3k``: colour, points, cost (scripts/game_logger.py:56-69)."""
    if card is None:
        return "[empty]"
    cost = "".join(f"{card.cost[c]}{_ABBR[i]}" for i, c in enumerate(("white", "blue", "green", "red", "black")) if card.cost.get(c, 0) > 0)
    return f"{card.color[0] if card.color != 'black' else 'k'}-{card.points}pt-{cost or 'free'}"


def describe_action(action: int, state) -> str:
    """The reference game logger's one-line description of ``action`` in ``state`` (scripts/game_logger.py:98-170;
    pinned string by string in tests/golden/logger_strings.json).  ``state``: a ``SplendorState`` mirror or a flat row."""
    from .engine.state import SplendorState, row_to_state

    s = state if isinstance(state, SplendorState) else row_to_state(state)
    a = int(action)
    if 0 <= a < 10:
        avail = [c for c in range(5) if s.bank[c] >= 1]
        if len(avail) >= 3:
            return "Take3: " + "".join(_ABBR[c] for c in TAKE3_COMBOS[a])
        if avail:  # the engine hands out what is left (engine/rules.py:45-58)
            return f"Take{len(avail)}: {''.join(_ABBR[c] for c in avail)} (reduced)"
    elif a < 15:
        return f"Take2: {_ABBR[a - 10] * 2}"
    elif a < 39:
        verb, k = ("Buy", a - 15) if a < 27 else ("Reserve", a - 27)
        return f"{verb}: T{1 + k // 4}S{1 + k % 4} {_card_text(s.board[1 + k // 4][k % 4])}"
    elif a < 42:
        return f"Reserve: T{a - 38} blind"
    elif a < 45:
        mine = s.players[s.to_play].reserved
        return f"BuyReserved: #{a - 41} {_card_text(mine[a - 42]) if a - 42 < len(mine) else '[empty]'}"
    return f"Action{a}"


@dataclass
class Trajectory:
    seed: int
    env_offset: int
    shuffle: str
    start_state: torch.Tensor    # uint8 [4, N, 16] packed hot rows at t = 0
    start_decks: torch.Tensor    # uint8 [N, 96]
    start_episode: torch.Tensor  # int32 [N]
    actions: torch.Tensor        # int32 [T, N]
    rewards: torch.Tensor        # float32 [T, N]
    terminated: torch.Tensor     # uint8 [T, N]

    @property
    def bytes_per_env_step(self) -> float:
        T, n = self.actions.shape
        fixed = self.start_state.numel() + self.start_decks.numel() + 4 * self.start_episode.numel()
        return 4.0 + fixed / (T * n)

    def save(self, path: str) -> None:
        torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}, path)

    @staticmethod
    def load(path: str) -> "Trajectory":
        return Trajectory(**torch.load(path))


def _snapshot(env: SplendorVecEnv):
    return env.state.clone(), env.decks.clone(), env.episode.clone()


def record_random(env: SplendorVecEnv, steps: int, seed: Optional[int] = None) -> Trajectory:
    """Random-legal rollout of `steps` lock-steps (one launch); only actions / rewards / terminations are kept.
    The trajectory records the env's OWN base seed (auto-reset deals are a function of it); ``seed`` is accepted for
    backwards compatibility and must agree with it."""
    if seed is not None and int(seed) != int(env._envs.seed_base):
        raise ValueError("record_random: seed differs from the env's base seed (the replay would deal other decks)")
    n, dev = env.n, env.device
    st, dk, ep = _snapshot(env)
    acts = torch.zeros((steps + 1, n), dtype=torch.int32, device=dev)
    rew = torch.zeros((steps, n), dtype=torch.float32, device=dev)
    term = torch.zeros((steps, n), dtype=torch.uint8, device=dev)
    env.sample_random_actions(out=acts[0])
    env.rollout_random(steps, acts[0], obs=None, mask=None, reward=rew, terminated=term, next_actions=acts)
    shuffle = "philox" if env.shuffle_mode == 1 else "mt19937"
    return Trajectory(int(env._envs.seed_base), int(env._envs.env_offset), shuffle, st, dk, ep, acts[:steps].clone(), rew, term)


def replay(traj: Trajectory, device="cuda", sink: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None,
           keep_obs: bool = False):
    """Re-execute a trajectory on the device.  Calls sink(t, obs[N,297], mask[N,45]) after every step (or collects the
    observations when keep_obs).  Returns (env, rewards, terminated, obs or None); rewards / terminations are
    re-derived, so comparing them with the log verifies the replay."""
    T, n = traj.actions.shape
    env = SplendorVecEnv(n, device=device, seed=traj.seed, shuffle=traj.shuffle, env_offset=traj.env_offset, autoreset=True)
    env.state.copy_(traj.start_state.to(env.device))
    env.decks.copy_(traj.start_decks.to(env.device))
    env.episode.copy_(traj.start_episode.to(env.device))
    env._is_reset = True
    rew = torch.zeros((T, n), dtype=torch.float32, device=env.device)
    term = torch.zeros((T, n), dtype=torch.uint8, device=env.device)
    obs_all = torch.zeros((T, n, 297), dtype=torch.int32, device=env.device) if keep_obs else None
    acts = traj.actions.to(env.device)
    for t in range(T):
        obs, _, _, _, _ = env.step(acts[t], out_obs=(obs_all[t] if keep_obs else None), out_reward=rew[t], out_terminated=term[t])
        if sink is not None:
            sink(t, obs, env.mask)
    return env, rew, term, obs_all
