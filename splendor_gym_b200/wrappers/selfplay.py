"""SelfPlayWrapper and random_opponent (splendor_gym/wrappers/selfplay.py:5-73) for the single-env facade, plus the
batched form for SplendorVecEnv.

Reward convention kept from the reference (:42-63): the agent is player 0; a game that ends on the agent's own move
pays the agent that move's reward; one that ends on the opponent's reply pays MINUS the opponent's reward -- which
makes a turn-limit draw worth +0.1 here, unlike the dual-step wrappers (pinned by tests/golden/wrappers.json).
The turn itself runs on the device (``_turn.play_turn`` -> ``SplendorVecEnv.dual_step(reward_mode="selfplay")``).
"""
from __future__ import annotations

import numpy as np

from ..envs._gym_compat import Wrapper
from ._turn import play_turn


def random_opponent(obs, info):
    """Uniform over the legal actions; 0 when there is none (wrappers/selfplay.py:66-73)."""
    mask = info.get("action_mask")
    if mask is None:
        return 0
    legal = np.flatnonzero(mask)
    if len(legal) == 0:
        return 0
    return int(np.random.choice(legal))


class _OpponentSeat(Wrapper):
    """What the three wrappers share: which opponent plays this episode (``opponent_supplier`` is asked once per reset,
    :21-25), and the opening moves when the env hands the first move to player 1 (:27-40)."""

    def __init__(self, env, opponent_policy, random_starts: bool = True, opponent_supplier=None):
        super().__init__(env)
        self.opponent_policy = opponent_policy
        self.random_starts = random_starts
        self.opponent_supplier = opponent_supplier
        self._opp_policy = opponent_policy

    def _opponent_moved(self) -> None:
        pass

    def reset(self, **kwargs):
        self._opp_policy = self.opponent_policy if self.opponent_supplier is None else self.opponent_supplier()
        obs, info = self.env.reset(**kwargs)
        if self.random_starts and info.get("to_play", 0) == 1:
            np.random.rand()  # the reference's coin (:29) decides only whether the first reply is played before or inside the loop
        while info.get("to_play", 0) == 1:  # (never true after SplendorEnv.reset: player 0 always starts)
            obs, _, term, trunc, info = self.env.step(self._opp_policy(obs, info))
            self._opponent_moved()
            if term or trunc:
                break
        return obs, info


class SelfPlayWrapper(_OpponentSeat):
    def step(self, action):
        turn = play_turn(self.env, action, self._opp_policy, "selfplay")
        if not turn.done and turn.opponent_action is None:
            raise RuntimeError(f"Invalid state: game not terminal but to_play={turn.info_agent.get('to_play', 'unknown')} (expected 1 for opponent)")
        return turn.obs, turn.agent_reward, turn.done, False, turn.info_final


def vec_selfplay_step(vec, agent_actions, opponent_policy):
    """SelfPlayWrapper.step for every env of a SplendorVecEnv: returns (obs, reward, terminated, truncated, info)
    with reward = r_agent if the agent's move ended the game, else -r_opponent if the opponent's did, else 0."""
    obs, agent_r, _, opp_r, done, info = vec.dual_step(agent_actions, opponent_policy, reward_mode="selfplay")
    return obs, agent_r, done, vec.truncated, info
