"""One self-play turn of the single-env facade, executed by the batched device path.

The three reference wrappers (wrappers/selfplay.py:42-63, wrappers/dual_step_native.py:90-193,
wrappers/dual_step_selfplay.py:83-158) all do the same thing per ``step``: the agent's move, then -- if the game goes
on -- one move of the opponent policy, and a reward convention on top.  Here that turn is ONE call of
``SplendorVecEnv(1).dual_step`` (two step-kernel launches + ``spl_dual_combine`` for the reward convention); the
wrappers only translate between the reference's per-env Python types and the device tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional

import numpy as np

from .. import _lib as L
from ..engine.encode import TOTAL_ACTIONS


@dataclass
class Turn:
    obs: np.ndarray                 # observation after the last move of the turn
    agent_reward: float             # per the reward convention asked for ("native" / "selfplay")
    opponent_reward: float
    done: bool
    ended_on: Optional[str]         # "agent_move" | "opponent_move" | None
    info_agent: Dict[str, Any]      # SplendorEnv.step's info after the agent's move
    info_final: Dict[str, Any]      # ... after the last move of the turn (== info_agent when the opponent did not move)
    opponent_action: Optional[int]


def play_turn(env, agent_action: int, opponent_policy: Callable, reward_mode: str) -> Turn:
    """``env``: the SplendorEnv facade (its ``.state`` mirror is pushed before and refreshed after the turn).
    Raises what ``SplendorEnv.step`` raises for the agent's move (envs/splendor_env.py:51-63)."""
    import torch

    assert env.state is not None, "Call reset() first"
    if env.state.game_over and env.state.to_play == 0:
        raise RuntimeError("Cannot call step() after episode termination. Call reset().")
    vec = env._engine()
    env._push()
    a = int(agent_action)
    seen: Dict[str, Any] = {}

    def opponent(obs_t, mask_t):
        # runs between the two launches: the device holds the position after the agent's move
        bits = int(vec.info_bits[0])
        ended = bool(vec._terminated[0])
        env._pull()
        seen["info"] = env._info(bits, ended, mask_t[0].cpu().numpy().astype(np.int8))
        seen["obs"] = obs_t[0].cpu().numpy().astype(np.int32)
        move = 0
        if not ended and not bits & (L.INFO_ILLEGAL | L.INFO_ERROR):  # otherwise dual_step leaves the env alone in phase 2
            move = int(opponent_policy(seen["obs"], seen["info"]))
            seen["move"] = move
        return torch.tensor([move if 0 <= move < TOTAL_ACTIONS else -1], dtype=torch.int32, device=vec.device)

    first = torch.tensor([a if 0 <= a < TOTAL_ACTIONS else -1], dtype=torch.int32, device=vec.device)
    obs, agent_r, _, opp_r, done, dinfo = vec.dual_step(first, opponent, reward_mode=reward_mode)
    bits1 = int(dinfo["info_bits_agent"][0])
    if bits1 & L.INFO_ERROR and not bits1 & L.INFO_NOLEGAL_DRAW:
        raise ValueError("Action out of bounds for action_space")
    info_agent = seen["info"]
    if "move" in seen:
        bits2 = int(dinfo["info_bits_opponent"][0])
        if bits2 & L.INFO_ERROR and not bits2 & L.INFO_NOLEGAL_DRAW:
            raise ValueError("Action out of bounds for action_space")
        env._pull()
        ended = bool(done[0])
        info_final = env._info(bits2, ended, vec.mask[0].cpu().numpy().astype(np.int8))
        return Turn(vec.obs[0].cpu().numpy().astype(np.int32), float(agent_r[0]), float(opp_r[0]), ended,
                    "opponent_move" if ended else None, info_agent, info_final, seen["move"])
    ended = bool(done[0])
    return Turn(seen["obs"], float(agent_r[0]), float(opp_r[0]), ended, "agent_move" if ended else None, info_agent, info_agent, None)
