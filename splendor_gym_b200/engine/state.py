"""Host-side state containers with the reference's field names (splendor_gym/engine/state.py:36-104) and the
conversion to / from the flat int32 row that the device kernels import / export
(include/splendor_b200.h, SPL_ROW_*).  These objects carry no rules: every rule evaluation goes to the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import tables

TOKEN_COLORS = ["white", "blue", "green", "red", "black", "gold"]
STANDARD_COLORS = TOKEN_COLORS[:-1]
COLOR_INDEX = {c: i for i, c in enumerate(TOKEN_COLORS)}
STANDARD_COLOR_INDEX = {c: i for i, c in enumerate(STANDARD_COLORS)}
HUMAN_TO_INTERNAL = {"diamond": "white", "sapphire": "blue", "emerald": "green", "ruby": "red", "onyx": "black"}
INTERNAL_TO_HUMAN = {v: k for k, v in HUMAN_TO_INTERNAL.items()}

ROW_LEN = 166
_DECK_OFF = {1: 76, 2: 116, 3: 146}


@dataclass(frozen=True)
class Card:
    id: int
    tier: int
    color: str
    points: int
    cost: Dict[str, int]

    def __hash__(self):
        return self.id


@dataclass(frozen=True)
class Noble:
    id: int  # 1000 + index, as in the reference (engine/state.py:161-174)
    requirements: Dict[str, int]
    points: int = 3

    def __hash__(self):
        return self.id


def _make_cards() -> List[Card]:
    out = []
    for i, (tier, col, pts, cost) in enumerate(tables.CARDS):
        out.append(Card(i, tier, STANDARD_COLORS[col], pts, {STANDARD_COLORS[k]: v for k, v in enumerate(cost) if v}))
    return out


def _make_nobles() -> List[Noble]:
    return [Noble(1000 + i, {STANDARD_COLORS[k]: v for k, v in enumerate(req) if v}, pts) for i, (req, pts) in enumerate(tables.NOBLES)]


CARDS: List[Card] = _make_cards()
NOBLES: List[Noble] = _make_nobles()


@dataclass
class PlayerState:
    tokens: List[int] = field(default_factory=lambda: [0] * 6)
    bonuses: List[int] = field(default_factory=lambda: [0] * 5)
    prestige: int = 0
    reserved: List[Card] = field(default_factory=list)
    revealed_reserved: List[bool] = field(default_factory=list)
    nobles: List[Noble] = field(default_factory=list)

    def can_afford(self, card: Card):
        """(affordable, per-colour cost after bonuses) -- engine/state.py:61-71.  Kept for API parity (the reference's
        bots and debugging helpers call it); the step path evaluates affordability on the device (spl_legal_mask)."""
        owed = [max(0, card.cost.get(c, 0) - self.bonuses[i]) for i, c in enumerate(STANDARD_COLORS)]
        gold_needed = sum(max(0, o - self.tokens[i]) for i, o in enumerate(owed))
        return self.tokens[COLOR_INDEX["gold"]] >= gold_needed, owed


@dataclass
class SplendorState:
    num_players: int
    bank: List[int]
    players: List[PlayerState]
    board: Dict[int, List[Optional[Card]]]
    decks: Dict[int, List[Card]]
    nobles: List[Optional[Noble]]
    to_play: int = 0
    turn_count: int = 1
    move_count: int = 0
    game_over: bool = False
    winner_index: Optional[int] = None
    turn_limit_reached: bool = False

    def copy(self) -> "SplendorState":
        return row_to_state(state_to_row(self))


def state_to_row(s: SplendorState) -> np.ndarray:
    row = np.full(ROW_LEN, -1, dtype=np.int32)
    row[0:6] = s.bank
    for p in range(2):
        pl = s.players[p]
        o = 6 + 23 * p
        row[o:o + 6] = pl.tokens
        row[o + 6:o + 11] = pl.bonuses
        row[o + 11] = pl.prestige
        nres = min(len(pl.reserved), 3)
        row[o + 12] = len(pl.reserved)
        for i in range(nres):
            row[o + 13 + i] = pl.reserved[i].id
            # a missing flag (the reference's own tests assign `reserved` without flags) reads as revealed
            row[o + 16 + i] = int(bool(pl.revealed_reserved[i])) if i < len(pl.revealed_reserved) else 1
        for i in range(nres, 3):
            row[o + 16 + i] = 0
        row[o + 19] = len(pl.nobles)
        for i, n in enumerate(pl.nobles[:3]):
            row[o + 20 + i] = n.id - 1000
    for t in (1, 2, 3):
        for k in range(4):
            c = s.board[t][k]
            row[52 + (t - 1) * 4 + k] = -1 if c is None else c.id
        row[64 + t - 1] = len(s.decks[t])
        for k, c in enumerate(s.decks[t]):
            row[_DECK_OFF[t] + k] = c.id
    for i in range(3):
        n = s.nobles[i] if i < len(s.nobles) else None
        row[67 + i] = -1 if n is None else n.id - 1000
    row[70] = s.to_play
    row[71] = s.turn_count
    row[72] = s.move_count
    row[73] = int(bool(s.game_over))
    row[74] = -1 if s.winner_index is None else int(s.winner_index)
    row[75] = int(bool(s.turn_limit_reached))
    return row


def row_to_state(row) -> SplendorState:
    r = [int(x) for x in row]
    players = []
    for p in range(2):
        o = 6 + 23 * p
        nres = r[o + 12]
        players.append(PlayerState(
            tokens=r[o:o + 6], bonuses=r[o + 6:o + 11], prestige=r[o + 11],
            reserved=[CARDS[r[o + 13 + i]] for i in range(min(nres, 3))],
            revealed_reserved=[bool(r[o + 16 + i]) for i in range(min(nres, 3))],
            nobles=[NOBLES[r[o + 20 + i]] for i in range(min(r[o + 19], 3))],
        ))
    board = {t: [None if r[52 + (t - 1) * 4 + k] < 0 else CARDS[r[52 + (t - 1) * 4 + k]] for k in range(4)] for t in (1, 2, 3)}
    decks = {t: [CARDS[r[_DECK_OFF[t] + k]] for k in range(r[64 + t - 1])] for t in (1, 2, 3)}
    nobles = [None if r[67 + i] < 0 else NOBLES[r[67 + i]] for i in range(3)]
    return SplendorState(
        num_players=2, bank=r[0:6], players=players, board=board, decks=decks, nobles=nobles, to_play=r[70],
        turn_count=r[71], move_count=r[72], game_over=bool(r[73]), winner_index=None if r[74] < 0 else r[74],
        turn_limit_reached=bool(r[75]),
    )
