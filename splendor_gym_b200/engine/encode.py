"""Action layout and observation size (splendor_gym/engine/encode.py:20-35,74).  The encoder itself is the
CUDA kernel (csrc/spl_core.cuh spl_encode_observation); `encode_observation` here calls it."""
from __future__ import annotations

import itertools

TAKE3_OFFSET, TAKE3_COUNT = 0, 10
TAKE2_OFFSET, TAKE2_COUNT = 10, 5
BUY_VISIBLE_OFFSET, BUY_VISIBLE_COUNT = 15, 12
RESERVE_VISIBLE_OFFSET, RESERVE_VISIBLE_COUNT = 27, 12
RESERVE_BLIND_OFFSET, RESERVE_BLIND_COUNT = 39, 3
BUY_RESERVED_OFFSET, BUY_RESERVED_COUNT = 42, 3
TOTAL_ACTIONS = 45
OBSERVATION_DIM = 297
TAKE3_COMBOS = list(itertools.combinations(range(5), 3))


def encode_take3_index(combo_index: int) -> int:
    return TAKE3_OFFSET + combo_index


def encode_take2_index(color_index: int) -> int:
    return TAKE2_OFFSET + color_index


def encode_buy_visible_index(tier: int, slot: int) -> int:
    return BUY_VISIBLE_OFFSET + (tier - 1) * 4 + slot


def encode_reserve_visible_index(tier: int, slot: int) -> int:
    return RESERVE_VISIBLE_OFFSET + (tier - 1) * 4 + slot


def encode_reserve_blind_index(tier: int) -> int:
    return RESERVE_BLIND_OFFSET + (tier - 1)


def encode_buy_reserved_index(slot: int) -> int:
    return BUY_RESERVED_OFFSET + slot


def encode_observation(state):
    from .rules import encode_observation as _enc

    return _enc(state)
