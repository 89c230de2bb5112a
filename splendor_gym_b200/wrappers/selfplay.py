"""SelfPlayWrapper and random_opponent (splendor_gym/wrappers/selfplay.py:20-73) for the single-env facade,
plus their batched form for SplendorVecEnv.

Semantics kept from the reference: the agent is player 0; after the agent's move the opponent policy is
queried once and its move applied; if that move ends the game the agent's reward is MINUS the opponent's
reward (:55-57) -- which makes a turn-limit draw worth +0.1 here, unlike the dual-step wrappers.
"""
from __future__ import annotations

import numpy as np

from ..envs._gym_compat import Wrapper


def random_opponent(obs, info):
    """Uniform over the legal actions; 0 when there is none (wrappers/selfplay.py:66-73)."""
    mask = info.get("action_mask")
    if mask is None:
        return 0
    legal = np.flatnonzero(mask)
    if len(legal) == 0:
        return 0
    return int(np.random.choice(legal))


class SelfPlayWrapper(Wrapper):
    def __init__(self, env, opponent_policy, random_starts: bool = True, opponent_supplier=None):
        super().__init__(env)
        self.opponent_policy = opponent_policy
        self.random_starts = random_starts
        self.opponent_supplier = opponent_supplier
        self._opp_policy = opponent_policy

    def _opponent_turns(self, obs, info):
        """Play opponent moves while it is player 1's turn (after reset this never loops: to_play == 0)."""
        while info.get("to_play", 0) == 1:
            obs, _, term, trunc, info = self.env.step(self._opp_policy(obs, info))
            if term or trunc:
                break
        return obs, info

    def reset(self, **kwargs):
        self._opp_policy = self.opponent_supplier() if self.opponent_supplier is not None else self.opponent_policy
        obs, info = self.env.reset(**kwargs)
        if self.random_starts and info.get("to_play", 0) == 1 and np.random.rand() < 0.5:
            obs, _, term, trunc, info = self.env.step(self._opp_policy(obs, info))
            if term or trunc:
                return obs, info
        return self._opponent_turns(obs, info)

    def step(self, action):
        obs, reward, term, trunc, info = self.env.step(action)
        if term or trunc:
            return obs, reward, term, trunc, info
        if info.get("to_play", 0) != 1:
            raise RuntimeError(f"Invalid state: game not terminal but to_play={info.get('to_play', 'unknown')} (expected 1 for opponent)")
        obs, opp_reward, term, trunc, info = self.env.step(self._opp_policy(obs, info))
        reward = -opp_reward if (term or trunc) else 0.0
        return obs, reward, term, trunc, info


def vec_selfplay_step(vec, agent_actions, opponent_policy):
    """SelfPlayWrapper.step for every env of a SplendorVecEnv: returns (obs, reward, terminated, truncated, info)
    with reward = r_agent if the agent's move ended the game, else -r_opponent if the opponent's did, else 0."""
    obs, agent_r, _, opp_r, done, info = vec.dual_step(agent_actions, opponent_policy, reward_mode="selfplay")
    return obs, agent_r, done, vec.truncated, info
