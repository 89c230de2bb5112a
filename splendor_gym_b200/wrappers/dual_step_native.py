"""DualStepNativeWrapper (splendor_gym/wrappers/dual_step_native.py:6-223): `dual_step(a)` = the agent's move,
the opponent policy's move, both rewards.  The batched equivalent is `SplendorVecEnv.dual_step`."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np

from ..envs._gym_compat import Wrapper
from .selfplay import random_opponent  # noqa: F401  (re-exported like the reference, :214-223)


class DualStepNativeWrapper(Wrapper):
    def __init__(self, env, opponent_policy: Callable, random_starts: bool = True, opponent_supplier: Optional[Callable] = None):
        super().__init__(env)
        self.opponent_policy = opponent_policy
        self.random_starts = random_starts
        self.opponent_supplier = opponent_supplier
        self._opp_policy = opponent_policy
        self.turn_count = 0
        self.total_agent_steps = 0
        self.total_opponent_steps = 0

    def reset(self, **kwargs):
        self._opp_policy = self.opponent_supplier() if self.opponent_supplier is not None else self.opponent_policy
        obs, info = self.env.reset(**kwargs)
        self.turn_count = self.total_agent_steps = self.total_opponent_steps = 0
        if self.random_starts and info.get("to_play", 0) == 1 and np.random.rand() < 0.5:
            obs, _, term, trunc, info = self.env.step(self._opp_policy(obs, info))
            self.total_opponent_steps += 1
            if term or trunc:
                return obs, info
        while info.get("to_play", 0) == 1:
            obs, _, term, trunc, info = self.env.step(self._opp_policy(obs, info))
            self.total_opponent_steps += 1
            if term or trunc:
                break
        return obs, info

    def step(self, action: int):
        agent_obs, agent_reward, _, _, done, info = self.dual_step(action)
        return agent_obs, agent_reward, done, False, info

    @staticmethod
    def _final_reward(info: Dict, player_id: int) -> float:
        fr = info.get("final_rewards")
        return fr[player_id] if fr is not None and player_id in fr else 0.0

    def dual_step(self, agent_action: int) -> Tuple[np.ndarray, float, np.ndarray, float, bool, Dict[str, Any]]:
        if getattr(self.env, "state", None) is None:
            raise RuntimeError("Cannot call dual_step() before reset()")
        if self.env.state.to_play != 0:
            raise ValueError("dual_step() requires agent (player 0) to move first")
        self.turn_count += 1
        self.total_agent_steps += 1
        obs1, r1, done1, trunc1, info1 = self.env.step(agent_action)
        turn_info = {"turn_count": self.turn_count, "agent_action": agent_action, "total_agent_steps": self.total_agent_steps,
                     "total_opponent_steps": self.total_opponent_steps, "phase": "agent_only"}
        turn_info.update(info1)
        if done1 or trunc1:
            opp_r = self._final_reward(info1, 1)
            turn_info.update({"opponent_action": None, "opponent_reward": opp_r, "turn_complete": True, "game_ended_on": "agent_move"})
            return obs1, r1, obs1, opp_r, True, turn_info
        if self.env.state.to_play != 1:
            raise ValueError(f"Expected opponent (player 1) to move after agent, got to_play={self.env.state.to_play}")
        opp_action = self._opp_policy(obs1, info1)
        self.total_opponent_steps += 1
        obs2, r2, done2, trunc2, info2 = self.env.step(opp_action)
        ended = done2 or trunc2
        agent_r = self._final_reward(info2, 0) if ended else 0.0
        turn_info.update(info2)
        turn_info.update({"opponent_action": opp_action, "opponent_reward": r2, "total_opponent_steps": self.total_opponent_steps,
                          "phase": "complete_turn", "turn_complete": True, "game_ended_on": "opponent_move" if ended else None})
        return obs2, agent_r, obs2, r2, ended, turn_info

    def get_wrapper_stats(self) -> Dict[str, Any]:
        return {"turn_count": self.turn_count, "total_agent_steps": self.total_agent_steps,
                "total_opponent_steps": self.total_opponent_steps,
                "avg_opponent_steps_per_turn": self.total_opponent_steps / max(1, self.turn_count),
                "wrapper_type": "DualStepNativeWrapper"}
