"""SplendorEnv -- the reference's single-environment Gymnasium API (splendor_gym/envs/splendor_env.py:23-130)
as a facade over a one-environment SplendorVecEnv.  Same constructor, spaces, return types, info keys and
error behaviour; every rule, mask and observation comes from the CUDA kernels.

`.state` is public in the reference and mutated by its tests and read by its wrappers
(tests/utils.py:25-54, wrappers/dual_step_native.py:108-114).  Here it is a host mirror object with the same
field names: it is pushed to the device before each step and refreshed after it, so in-place edits such as
``env.state.bank[:] = [0]*6`` take effect exactly as they do on the reference.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import numpy as np

from .. import _lib as L
from ..engine.encode import OBSERVATION_DIM, TOTAL_ACTIONS
from ..engine.state import SplendorState, row_to_state, state_to_row
from ._gym_compat import Env, spaces


class SplendorEnv(Env):
    metadata = {"render_modes": ["human"], "name": "Splendor-v0"}

    def __init__(self, num_players: int = 2, render_mode: str | None = None, seed: int | None = None, device="cuda"):
        super().__init__()
        if num_players != 2:
            raise NotImplementedError("Current env supports 2 players only.")
        self.num_players = num_players
        self.render_mode = render_mode
        self.action_space = spaces.Discrete(TOTAL_ACTIONS)
        self.observation_space = spaces.Box(low=0, high=50, shape=(OBSERVATION_DIM,), dtype=np.int32)
        self.state: SplendorState | None = None
        self.current_player: int = 0
        self._device = device
        self._vec = None

    # ------------------------------------------------------------------ device plumbing
    def _engine(self):
        if self._vec is None:
            from ..vec_env import SplendorVecEnv

            self._vec = SplendorVecEnv(1, device=self._device, shuffle="mt19937", autoreset=False)
        return self._vec

    def _push(self):
        import torch

        self._engine().import_state(torch.from_numpy(state_to_row(self.state)[None, :]))

    def _pull(self):
        self.state = row_to_state(self._engine().export_state()[0].cpu().numpy())

    # ------------------------------------------------------------------ gym API
    def reset(self, *, seed: int | None = None, options: Dict[str, Any] | None = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        import torch

        super().reset(seed=seed)
        engine_seed = int(self.np_random.integers(0, 2**31 - 1))  # envs/splendor_env.py:43
        vec = self._engine()
        obs, info = vec.reset(seeds=torch.tensor([engine_seed], dtype=torch.int64))
        self._pull()
        self.current_player = self.state.to_play
        return obs[0].cpu().numpy().astype(np.int32), {"action_mask": info["action_mask"][0].cpu().numpy().astype(np.int8),
                                                        "to_play": int(self.state.to_play)}

    def step(self, action: int) -> Tuple[np.ndarray, float, bool, bool, Dict[str, Any]]:
        import torch

        assert self.state is not None, "Call reset() first"
        if self.state.game_over and self.state.to_play == 0:
            raise RuntimeError("Cannot call step() after episode termination. Call reset().")
        vec = self._engine()
        self._push()
        a = int(action)
        in_range = 0 <= a < TOTAL_ACTIONS
        vec.step(torch.tensor([a if in_range else -1], dtype=torch.int32), autoreset=False)
        bits = int(vec.info_bits[0])
        if bits & L.INFO_ERROR and not (bits & L.INFO_NOLEGAL_DRAW):
            # the no-legal-move draw is checked BEFORE the bounds check in the reference (:55-63)
            raise ValueError("Action out of bounds for action_space")
        self._pull()
        obs = vec.obs[0].cpu().numpy().astype(np.int32)
        mask = vec.mask[0].cpu().numpy().astype(np.int8)
        reward = float(vec.reward[0])
        terminated = bool(vec.terminated[0])
        return obs, reward, terminated, False, self._info(bits, terminated, mask)

    def _info(self, bits: int, terminated: bool, mask: np.ndarray) -> Dict[str, Any]:
        """The reference's info dict (envs/splendor_env.py:61,66,81-88) from the kernel's info bits; ``self.state`` is current."""
        info: Dict[str, Any] = {"action_mask": mask, "to_play": int(self.state.to_play)}
        if bits & L.INFO_NOLEGAL_DRAW:
            info["draw"] = True
        if bits & L.INFO_ILLEGAL:
            info["illegal_action"] = True
        if terminated and bits & L.INFO_TURN_LIMIT:
            info["turn_limit"] = True
        if terminated and not (bits & L.INFO_NOLEGAL_DRAW):
            info["final_rewards"] = self.get_final_rewards()
        return info

    def get_final_rewards(self) -> Dict[int, float]:
        """envs/splendor_env.py:92-115."""
        if not (self.state.game_over and self.state.to_play == 0):
            raise RuntimeError("Cannot get final rewards for non-terminal state")
        w = self.state.winner_index
        if w is None:
            r = -0.1 if self.state.turn_limit_reached else 0.0
            return {p: r for p in range(self.num_players)}
        return {p: (1.0 if w == p else -1.0) for p in range(self.num_players)}

    def render(self):
        if self.render_mode not in ("human", None):
            return
        assert self.state is not None
        s = self.state
        print(f"turn {s.turn_count} move {s.move_count} to_play {s.to_play} bank {s.bank}")
        for i, p in enumerate(s.players):
            print(f"  P{i}: tokens {p.tokens} bonuses {p.bonuses} prestige {p.prestige} reserved {[c.id for c in p.reserved]}")
        for t in (1, 2, 3):
            print(f"  tier {t}: {[None if c is None else c.id for c in s.board[t]]} deck {len(s.decks[t])}")


def make(num_players: int = 2, render_mode: str | None = None, seed: int | None = None) -> SplendorEnv:
    return SplendorEnv(num_players=num_players, render_mode=render_mode, seed=seed)
