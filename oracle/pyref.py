"""Loader for the LIVE Python reference (test infrastructure only).

This module exists so that the golden-vector generator (oracle/gen_golden.py) and the
in-container cross-check tests can execute the unmodified reference from
``/root/reference``.  It is never imported by the product package, by ``bench.py``'s GPU
arm or by anything that runs on the GPU box (``/root/reference`` does not exist there):
call ``available()`` first.

The reference package imports ``gymnasium`` (splendor_gym/__init__.py:1 ->
envs/splendor_env.py:3), which is not installed in this image.  A minimal stand-in that
provides only what the reference touches is registered in ``sys.modules`` before import:
``gymnasium.Env`` (PCG64 seeding exactly like gymnasium.utils.seeding.np_random),
``gymnasium.Wrapper`` and ``gymnasium.spaces.{Discrete,Box}``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SPLENDOR_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "splendor_gym", "engine", "rules.py"))


def _install_gymnasium_stub() -> str:
    """Register a stand-in ``gymnasium`` if the real one is missing. Returns which is in use."""
    try:
        import gymnasium  # noqa: F401

        return "real"
    except Exception:
        pass

    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Env:  # what envs/splendor_env.py:23-43 needs
        metadata: dict = {}
        _np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            return self._np_random

        def close(self):
            pass

    class Wrapper:  # what wrappers/*.py need
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def step(self, a):
            return self.env.step(a)

    class Discrete:
        def __init__(self, n):
            self.n = int(n)

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    gym.Env = Env
    gym.Wrapper = Wrapper
    gym.spaces = spaces
    spaces.Discrete = Discrete
    spaces.Box = Box
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    return "stub"


def load():
    """Import the reference package; returns a namespace with the modules used by the oracle tools."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    which = _install_gymnasium_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import splendor_gym  # noqa: F401
    from splendor_gym.engine import encode, rules, state
    from splendor_gym.envs import splendor_env
    from splendor_gym.wrappers import dual_step_native, selfplay

    ns = types.SimpleNamespace(
        gymnasium=which,
        state=state,
        rules=rules,
        encode=encode,
        env=splendor_env,
        selfplay=selfplay,
        dual_step_native=dual_step_native,
    )
    _loaded = ns
    return ns


# ---------------------------------------------------------------------------------------
# Reference dataclass  ->  canonical flat int32 "state row" (layout in include/splendor_b200.h,
# SPL_ROW_*).  Used to compare full state after every step.
# ---------------------------------------------------------------------------------------
ROW_LEN = 166


def state_to_row(s) -> np.ndarray:
    row = np.full(ROW_LEN, -1, dtype=np.int32)
    row[0:6] = s.bank
    for p in range(2):
        pl = s.players[p]
        o = 6 + 23 * p
        row[o : o + 6] = pl.tokens
        row[o + 6 : o + 11] = pl.bonuses
        row[o + 11] = pl.prestige
        row[o + 12] = len(pl.reserved)
        for i, c in enumerate(pl.reserved[:3]):
            row[o + 13 + i] = c.id
        for i in range(3):
            row[o + 16 + i] = int(bool(pl.revealed_reserved[i])) if i < len(pl.revealed_reserved) else 0
        row[o + 19] = len(pl.nobles)
        for i, n in enumerate(pl.nobles[:3]):
            row[o + 20 + i] = n.id - 1000
    for t in (1, 2, 3):
        for k in range(4):
            c = s.board[t][k]
            row[52 + (t - 1) * 4 + k] = -1 if c is None else c.id
    for t in (1, 2, 3):
        row[64 + t - 1] = len(s.decks[t])
    for i in range(3):
        n = s.nobles[i] if i < len(s.nobles) else None
        row[67 + i] = -1 if n is None else n.id - 1000
    row[70] = s.to_play
    row[71] = s.turn_count
    row[72] = s.move_count
    row[73] = int(bool(s.game_over))
    row[74] = -1 if s.winner_index is None else int(s.winner_index)
    row[75] = int(bool(s.turn_limit_reached))
    off = {1: 76, 2: 116, 3: 146}
    for t in (1, 2, 3):
        for k, c in enumerate(s.decks[t]):
            row[off[t] + k] = c.id
    return row


def row_to_state(row):
    """Inverse of state_to_row: builds a reference SplendorState (needs the live reference)."""
    ns = load()
    st = ns.state
    cards_by_tier = st._load_cards_from_json()
    by_id = {c.id: c for t in (1, 2, 3) for c in cards_by_tier[t]}
    nobles = {n.id - 1000: n for n in st._load_nobles_from_json()}
    row = [int(x) for x in row]
    players = []
    for p in range(2):
        o = 6 + 23 * p
        nres = row[o + 12]
        pl = st.PlayerState(
            tokens=row[o : o + 6],
            bonuses=row[o + 6 : o + 11],
            prestige=row[o + 11],
            reserved=[by_id[row[o + 13 + i]] for i in range(nres)],
            revealed_reserved=[bool(row[o + 16 + i]) for i in range(nres)],
            nobles=[nobles[row[o + 20 + i]] for i in range(row[o + 19])],
        )
        players.append(pl)
    board = {t: [None if row[52 + (t - 1) * 4 + k] < 0 else by_id[row[52 + (t - 1) * 4 + k]] for k in range(4)] for t in (1, 2, 3)}
    off = {1: 76, 2: 116, 3: 146}
    decks = {t: [by_id[row[off[t] + k]] for k in range(row[64 + t - 1])] for t in (1, 2, 3)}
    nob = [None if row[67 + i] < 0 else nobles[row[67 + i]] for i in range(3)]
    return st.SplendorState(
        num_players=2,
        bank=row[0:6],
        players=players,
        board=board,
        decks=decks,
        nobles=nob,
        to_play=row[70],
        turn_count=row[71],
        move_count=row[72],
        game_over=bool(row[73]),
        winner_index=None if row[74] < 0 else row[74],
        turn_limit_reached=bool(row[75]),
    )
