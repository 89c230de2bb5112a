"""The reference's functional engine API (splendor_gym/engine/__init__.py:1-13) on top of the CUDA kernels.

    initial_state(num_players=2, seed=0)   engine/state.py:181-211  -> spl_reset (MT19937 mode, bit-exact)
    legal_moves(state) -> List[int]        engine/rules.py:40-93    -> spl_observe
    apply_action(state, action) -> state   engine/rules.py:196-287  -> spl_step
    is_terminal / winner                   engine/rules.py:306-312
    encode_observation(state)              engine/encode.py:124-187 -> spl_observe

Each call moves one state row to a private one-environment SplendorVecEnv on the current CUDA device and
back; it exists for API parity (tests, debugging, single games), not for throughput -- batched work should
use SplendorVecEnv directly.  No rule is evaluated on the host.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .state import SplendorState, row_to_state, state_to_row

_ctx = None


def _engine():
    global _ctx
    if _ctx is None:
        import torch

        from ..vec_env import SplendorVecEnv

        if not torch.cuda.is_available():
            from .._lib import SplendorB200Error

            raise SplendorB200Error("the Splendor engine runs on a CUDA device; there is no CPU fallback")
        _ctx = SplendorVecEnv(1, device="cuda", shuffle="mt19937", autoreset=False)
    return _ctx


def _load(state: SplendorState):
    import torch

    e = _engine()
    e.import_state(torch.from_numpy(state_to_row(state)[None, :]))
    return e


def initial_state(num_players: int = 2, seed: int = 0) -> SplendorState:
    if num_players != 2:
        raise NotImplementedError("Current engine supports 2 players only.")
    import torch

    e = _engine()
    e.reset(seeds=torch.tensor([int(seed)], dtype=torch.int64))
    return row_to_state(e.export_state()[0].cpu().numpy())


def legal_moves(state: SplendorState) -> List[int]:
    """45-entry 0/1 list for the player to move.  Unlike the env's info["action_mask"], the reference's
    legal_moves does not zero the mask of a finished game; neither does this."""
    e = _load(state)
    if state.game_over and state.to_play == 0:  # spl_observe zeroes terminal masks (envs/splendor_env.py:81)
        s2 = row_to_state(state_to_row(state))
        s2.game_over = False
        e = _load(s2)
    _, mask = e.observe()
    return [int(x) for x in mask[0].cpu().numpy()]


def encode_observation(state: SplendorState) -> np.ndarray:
    e = _load(state)
    obs, _ = e.observe()
    return obs[0].cpu().numpy().astype(np.int32)


def apply_action(state: SplendorState, action: int) -> SplendorState:
    """Pure: returns a new state.  The action must be legal (the reference asserts on an empty slot and
    otherwise trusts the caller, engine/rules.py:216-257); an out-of-range index raises ValueError (:256-257)."""
    import torch

    from .. import _lib as L

    if not (0 <= int(action) < 45):
        raise ValueError("Invalid action index")
    e = _load(state)
    e.step(torch.tensor([int(action)], dtype=torch.int32), autoreset=False)
    bits = int(e.info_bits[0])
    if bits & (L.INFO_ILLEGAL | L.INFO_ERROR):
        raise ValueError(f"apply_action: action {action} is not legal in this state")
    return row_to_state(e.export_state()[0].cpu().numpy())


def is_terminal(state: SplendorState) -> bool:
    return bool(state.game_over and state.to_play == 0)


def winner(state: SplendorState) -> Optional[int]:
    return state.winner_index
