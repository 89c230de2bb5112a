"""DualStepNativeWrapper (splendor_gym/wrappers/dual_step_native.py:6-223): ``dual_step(a)`` = the agent's move, the
opponent policy's reply, both players' rewards.  An adapter: the turn is one ``SplendorVecEnv(1).dual_step`` on the
device (``_turn.play_turn``); the batched form for many envs is ``SplendorVecEnv.dual_step`` itself."""
from __future__ import annotations

from typing import Any, Dict, Tuple

import numpy as np

from ._turn import play_turn
from .selfplay import _OpponentSeat, random_opponent  # noqa: F401  (random_opponent is re-exported like the reference, :214-223)


class DualStepNativeWrapper(_OpponentSeat):
    def __init__(self, env, opponent_policy, random_starts: bool = True, opponent_supplier=None):
        super().__init__(env, opponent_policy, random_starts, opponent_supplier)
        self.turn_count = self.total_agent_steps = self.total_opponent_steps = 0

    def _opponent_moved(self) -> None:
        self.total_opponent_steps += 1

    def reset(self, **kwargs):
        self.turn_count = self.total_agent_steps = self.total_opponent_steps = 0
        return super().reset(**kwargs)

    def step(self, action: int):
        obs, reward, _, _, done, info = self.dual_step(action)
        return obs, reward, done, False, info

    def dual_step(self, agent_action: int) -> Tuple[np.ndarray, float, np.ndarray, float, bool, Dict[str, Any]]:
        """-> (agent_obs, agent_reward, opponent_obs, opponent_reward, done, info); both observations are the position
        after the turn (:182-191); the agent's reward of a finished game is final_rewards[0] (:159-161)."""
        state = getattr(self.env, "state", None)
        if state is None:
            raise RuntimeError("Cannot call dual_step() before reset()")
        if state.to_play != 0:
            raise ValueError("dual_step() requires agent (player 0) to move first")
        turn = play_turn(self.env, agent_action, self._opp_policy, "native")
        if not turn.done and turn.opponent_action is None:
            raise ValueError(f"Expected opponent (player 1) to move after agent, got to_play={turn.info_agent.get('to_play')}")
        self.turn_count += 1
        self.total_agent_steps += 1
        replied = turn.opponent_action is not None
        self.total_opponent_steps += int(replied)
        info: Dict[str, Any] = {"turn_count": self.turn_count, "agent_action": agent_action, "total_agent_steps": self.total_agent_steps}
        info.update(turn.info_agent)
        info.update(turn.info_final)
        info.update(total_opponent_steps=self.total_opponent_steps, phase="complete_turn" if replied else "agent_only",
                    opponent_action=turn.opponent_action, opponent_reward=turn.opponent_reward, turn_complete=True, game_ended_on=turn.ended_on)
        return turn.obs, turn.agent_reward, turn.obs, turn.opponent_reward, turn.done, info

    def get_wrapper_stats(self) -> Dict[str, Any]:
        return dict(turn_count=self.turn_count, total_agent_steps=self.total_agent_steps, total_opponent_steps=self.total_opponent_steps,
                    avg_opponent_steps_per_turn=self.total_opponent_steps / max(1, self.turn_count), wrapper_type="DualStepNativeWrapper")
