"""DualStepSelfPlayWrapper (splendor_gym/wrappers/dual_step_selfplay.py:6-186): a SelfPlayWrapper-shaped ``step`` whose
reward for a game that ends on the opponent's reply is ``final_rewards[0]`` (:138-152) instead of the negated opponent
reward -- a turn-limit draw is -0.1 here and +0.1 under SelfPlayWrapper.  Same device turn as the other two wrappers."""
from __future__ import annotations

from typing import Any, Dict

from ._turn import play_turn
from .selfplay import _OpponentSeat, random_opponent  # noqa: F401  (:177-186)


class DualStepSelfPlayWrapper(_OpponentSeat):
    def __init__(self, env, opponent_policy, random_starts: bool = True, opponent_supplier=None):
        super().__init__(env, opponent_policy, random_starts, opponent_supplier)
        self.turn_count = self.total_agent_actions = self.total_opponent_actions = 0

    def _opponent_moved(self) -> None:
        self.total_opponent_actions += 1

    def reset(self, **kwargs):
        self.turn_count = self.total_agent_actions = self.total_opponent_actions = 0
        return super().reset(**kwargs)

    def step(self, agent_action: int):
        turn = play_turn(self.env, agent_action, self._opp_policy, "native")
        if not turn.done and turn.opponent_action is None:
            raise RuntimeError(f"Invalid state after agent move: to_play={turn.info_agent.get('to_play', 'unknown')}, "
                               "expected 1 for opponent. Game state may be corrupted.")
        self.turn_count += 1
        self.total_agent_actions += 1
        info: Dict[str, Any] = {"turn_count": self.turn_count, "agent_action": agent_action, "total_agent_actions": self.total_agent_actions,
                                "total_opponent_actions": self.total_opponent_actions, "phase": "agent_only"}
        info.update(turn.info_agent)
        if turn.opponent_action is not None:
            self.total_opponent_actions += 1
            info.update(turn.info_final)
            info.update(opponent_action=turn.opponent_action, opponent_reward=turn.opponent_reward,
                        total_opponent_actions=self.total_opponent_actions, phase="complete_turn")
        info["turn_complete"] = True
        if turn.ended_on is not None:
            info["game_ended_on"] = turn.ended_on
        return turn.obs, turn.agent_reward, turn.done, False, info

    def get_wrapper_stats(self) -> Dict[str, Any]:
        return dict(turn_count=self.turn_count, total_agent_actions=self.total_agent_actions, total_opponent_actions=self.total_opponent_actions,
                    avg_opponent_actions_per_turn=self.total_opponent_actions / max(1, self.turn_count), wrapper_type="DualStepSelfPlayWrapper")
