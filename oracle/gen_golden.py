#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING THE UNMODIFIED REFERENCE (/root/reference).

Run in the build container only:   python oracle/gen_golden.py
The fixtures pin the C oracle (oracle/splendor_oracle.c) and, through it, the CUDA path.  They are
produced exclusively with reference code + CPython's `random` + numpy's PCG64 -- nothing from this
repository's engine is involved in computing an expected value.

Files written:
  mt19937.json        CPython random.Random(seed) raw 32-bit outputs (seeds incl. >32-bit ones)
  token_return.json   sha256 of the top-3-bit streams of all 8,910 in-domain auto_return_tokens seeds
  initial_states.json initial_state(seed) flat rows + obs sha (incl. the SURVEY.md section 8c fingerprints)
  env_seeding.json    SplendorEnv.reset(seed) -> engine seed / board (gymnasium PCG64 path)
  games.json          full games under 3 deterministic policies: actions + per-step digest of
                      (obs, mask, state row, reward, terminated, info bits) + a few full vectors
  edge_cases.json     hand-built states mirroring the reference's own tests (tests/utils.py style
                      mutation of env.state): full input row, action, and every output
  bots.json           (obs, mask) states + the decisions of the scripted opponents of scripts/eval_suite.py
  games_digest.json   10,000 reference games under the LCG policy: moves, winner and ONE sha256 per game over every step's
                      (observation, mask, reward, terminated, info bits)  [python oracle/gen_golden.py games_digest: ~1 min on 8 cores]
  autoreset_stream.json  reference envs auto-reset four times from reset(seed=s): engine seeds, episode start rows, digests
  logger_strings.json decode_action strings of scripts/game_logger.py and PlayerState.can_afford results on game positions
  wrappers.json       SelfPlayWrapper / DualStepNativeWrapper / DualStepSelfPlayWrapper turns (incl. turn-limit draws):
                      rewards, done flags, opponent moves, observation digests per wrapper step
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyref  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ns = pyref.load()
R, ENC, ST, ENV = ns.rules, ns.encode, ns.state, ns.env


def info_bits(info: dict, state) -> int:
    b = 0
    if info.get("illegal_action"):
        b |= 1
    if info.get("draw"):
        b |= 2
    if info.get("turn_limit"):
        b |= 4
    if R.is_terminal(state):
        b |= 8
        if "final_rewards" in info:
            w = state.winner_index
            b |= (0 if w is None else w + 1) << 4
    return b


def digest(obs, mask, row, reward, term, bits) -> str:
    h = hashlib.sha256()
    h.update(np.asarray(obs, np.int32).tobytes())
    h.update(np.asarray(mask, np.int8).tobytes())
    h.update(np.asarray(row, np.int32).tobytes())
    h.update(struct.pack("<fBB", float(reward), int(bool(term)), bits))
    return h.hexdigest()[:12]


def env_from_state(state):
    env = ENV.SplendorEnv()
    env.reset(seed=0)
    env.state = state
    return env


def dump(name, obj):
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


# ----------------------------------------------------------------------------- mt19937.json
def gen_mt():
    seeds = [0, 1, 42, 19650218, 2**31 - 2, 2**32 - 1, 2**32, 2**32 + 1, (99 * 1315423911) ^ 2654435761 ^ (13 * 97531) ^ (14 * 31337),
             2**40 + 7, 2**63 + 11]
    out = []
    for s in seeds:
        r = random.Random(s)
        outs = [r.getrandbits(32) for _ in range(1000)]  # crosses the 624-word regeneration boundary
        out.append({"seed": str(s), "first": outs[:8], "sha_1000": hashlib.sha256(struct.pack("<1000I", *outs)).hexdigest()})
    dump("mt19937.json", out)


# ----------------------------------------------------------------------------- token_return.json
def gen_token_return():
    """auto_return_tokens seed domain in legitimate play (engine/rules.py:160-165): turn_count 1..99,
    to_play 0..1, hand sum 11..13, bank sum 0..14.  getrandbits(k<=3) only uses the top 3 bits of each
    32-bit output; 21 outputs are stored per seed."""
    h = hashlib.sha256()
    samples = {}
    for turn in range(1, 100):
        for tp in range(2):
            for hand in range(11, 14):
                for bank in range(0, 15):
                    seed = (turn * 1315423911) ^ (tp * 2654435761) ^ (hand * 97531) ^ (bank * 31337)
                    r = random.Random(seed)
                    v = 0
                    for j in range(21):
                        v |= (r.getrandbits(32) >> 29) << (3 * j)
                    h.update(struct.pack("<Q", v))
                    if (turn, tp, hand, bank) in ((1, 0, 11, 0), (7, 1, 12, 9), (50, 0, 13, 14), (99, 1, 13, 14)):
                        samples[f"{turn},{tp},{hand},{bank}"] = str(v)
    dump("token_return.json", {"order": "turn,to_play,hand,bank (bank fastest)", "sha256": h.hexdigest(), "samples": samples})


# ----------------------------------------------------------------------------- initial_states.json
def gen_initial():
    out = []
    for seed in [0, 42, 123456789, 1, 2, 3, 7, 999, 2**31 - 2, 1826701614, 191664963, 33158374]:
        s = R.initial_state(seed=seed)
        obs = ENC.encode_observation(s)
        out.append({"seed": seed, "row": pyref.state_to_row(s).tolist(),
                    "obs_sha16": hashlib.sha256(obs.tobytes()).hexdigest()[:16],
                    "mask": R.legal_moves(s)})
    dump("initial_states.json", out)


# ----------------------------------------------------------------------------- env_seeding.json
def gen_env_seeding():
    out = []
    for seed in [0, 42, 123, 7]:
        env = ENV.SplendorEnv()
        obs, info = env.reset(seed=seed)
        eng = int(np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))).integers(0, 2**31 - 1))
        assert np.array_equal(pyref.state_to_row(env.state), pyref.state_to_row(R.initial_state(seed=eng)))
        # a second, unseeded reset keeps drawing from the same stream (ppo_splendor.py:246-247)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        e1 = int(g.integers(0, 2**31 - 1))
        e2 = int(g.integers(0, 2**31 - 1))
        env.reset()
        assert np.array_equal(pyref.state_to_row(env.state), pyref.state_to_row(R.initial_state(seed=e2)))
        out.append({"seed": seed, "engine_seed": e1, "engine_seed_2nd_reset": e2,
                    "gymnasium": ns.gymnasium, "numpy": np.__version__,
                    "tier1_board": [c.id for c in R.initial_state(seed=e1).board[1]]})
    dump("env_seeding.json", out)


# ----------------------------------------------------------------------------- games.json
def play(seed: int, policy: str):
    state = R.initial_state(seed=seed)
    env = env_from_state(state)
    x = (seed * 2654435761) % 2**32
    actions, digests, full = [], [], {}
    t = 0
    while True:
        mask = R.legal_moves(env.state)
        legal = [i for i, v in enumerate(mask) if v]
        x = (1664525 * x + 1013904223) % 2**32
        if not legal:
            a = 0
        elif policy == "first":
            a = legal[0]
        elif policy == "lcg":
            a = legal[(x >> 16) % len(legal)]
        elif policy == "lcg_illegal":
            a = (x >> 8) % 45 if t % 7 == 3 else legal[(x >> 16) % len(legal)]
        else:
            raise ValueError(policy)
        obs, r, term, trunc, info = env.step(a)
        row = pyref.state_to_row(env.state)
        bits = info_bits(info, env.state)
        actions.append(a)
        digests.append(digest(obs, info["action_mask"], row, r, term, bits))
        if t in (0, 10, 40) or term:
            full[str(t)] = {"obs": obs.tolist(), "mask": info["action_mask"].tolist(), "row": row.tolist(),
                            "reward": float(r), "terminated": bool(term), "info": bits}
        t += 1
        if term or t >= 400:
            break
    s = env.state
    return {"seed": seed, "policy": policy, "actions": actions, "digests": digests, "full": full,
            "moves": s.move_count, "winner": s.winner_index, "prestige": [p.prestige for p in s.players]}


def gen_games():
    games = []
    for seed in (0, 1, 999):
        games.append(play(seed, "first"))
        games.append(play(seed, "lcg"))
    for seed in range(100, 118):
        games.append(play(seed, "lcg"))
    for seed in range(200, 212):
        games.append(play(seed, "lcg_illegal"))
    # SURVEY.md section 8c fingerprints (first-legal / LCG policies): moves, winner, prestige
    g = {(x["seed"], x["policy"]): x for x in games}
    assert (g[(0, "first")]["moves"], g[(0, "first")]["winner"], g[(0, "first")]["prestige"]) == (116, 0, [17, 16])
    assert (g[(1, "lcg")]["moves"], g[(1, "lcg")]["winner"], g[(1, "lcg")]["prestige"]) == (62, 1, [6, 17])
    dump("games.json", games)


# ----------------------------------------------------------------------------- edge_cases.json
def record(name, state, action, cite):
    """Run SplendorEnv.step on a (possibly hand-mutated) state; store input row and every output."""
    env = env_from_state(state)
    row_in = pyref.state_to_row(env.state)
    mask_in = R.legal_moves(env.state)
    try:
        obs, r, term, trunc, info = env.step(action)
    except (RuntimeError, ValueError) as e:
        return {"name": name, "cite": cite, "row_in": row_in.tolist(), "mask_in": mask_in, "action": action,
                "raises": type(e).__name__}
    return {"name": name, "cite": cite, "row_in": row_in.tolist(), "mask_in": mask_in, "action": action,
            "row_out": pyref.state_to_row(env.state).tolist(), "obs": obs.tolist(),
            "mask": info["action_mask"].tolist(), "reward": float(r), "terminated": bool(term),
            "info": info_bits(info, env.state),
            "final_rewards": [info["final_rewards"][0], info["final_rewards"][1]] if "final_rewards" in info else None}


def cards_by_id():
    c = ST._load_cards_from_json()
    return {x.id: x for t in (1, 2, 3) for x in c[t]}


def gen_edges():
    C = cards_by_id()
    out = []

    # tests/test_rules.py:37-43 -- out-of-domain hand [5,5,5,5,5,0] then any action -> return to 10
    s = R.initial_state(seed=0)
    s.players[0].tokens = [5, 5, 5, 5, 5, 0]
    out.append(record("token_limit_25_tokens", s, 0, "tests/test_rules.py:37-43"))
    # tests/test_afford_nobles_obs.py:58-71 -- hand [3,3,3,3,3,0]
    s = R.initial_state(seed=int(np.random.Generator(np.random.PCG64(np.random.SeedSequence(7))).integers(0, 2**31 - 1)))
    s.players[0].tokens = [3, 3, 3, 3, 3, 0]
    out.append(record("token_return_15_tokens", s, [i for i, v in enumerate(R.legal_moves(s)) if v][0],
                      "tests/test_afford_nobles_obs.py:58-71"))
    # gold-only overflow: hand of 9 gold + 2 white, take 3 -> must give back non-gold first then gold
    s = R.initial_state(seed=5)
    s.players[0].tokens = [0, 0, 0, 0, 0, 10]
    out.append(record("token_return_gold_last_resort", s, 0, "engine/rules.py:169-184"))
    # tests/test_draw_rule.py:7-24
    s = R.initial_state(seed=0)
    s.bank[:] = [0, 0, 0, 0, 0, 0]
    s.players[0].tokens[:] = [10, 0, 0, 0, 0, 0]
    s.players[0].reserved = s.decks[1][:3]
    s.players[0].revealed_reserved = [True, True, True]
    for t in (1, 2, 3):
        s.board[t] = [None, None, None, None]
    out.append(record("no_legal_move_draw", s, 0, "tests/test_draw_rule.py:7-24; envs/splendor_env.py:55-61"))
    # same, but on player 1's turn with game_over already set by player 0 (state forced)
    s = R.initial_state(seed=3)
    s = R.apply_action(s, 0)
    s.bank[:] = [0, 0, 0, 0, 0, 0]
    s.players[1].tokens[:] = [0, 0, 0, 0, 0, 0]
    s.players[1].reserved = s.decks[2][:3]
    s.players[1].revealed_reserved = [False, True, False]
    for t in (1, 2, 3):
        s.board[t] = [None, None, None, None]
    out.append(record("no_legal_move_draw_p1", s, 7, "envs/splendor_env.py:55-61"))
    # tests/test_take_reduced_colors.py:7-21 and :24-36
    s = R.initial_state(seed=123)
    s.bank[:] = [1, 0, 2, 0, 0, 0]
    out.append(record("take3_two_colours", s, 0, "tests/test_take_reduced_colors.py:7-21"))
    s = R.initial_state(seed=123)
    s.bank[:] = [1, 0, 2, 0, 0, 0]
    out.append(record("take3_two_colours_illegal_combo", s, 1, "tests/test_take_reduced_colors.py:7-21"))
    s = R.initial_state(seed=123)
    s.bank[:] = [0, 0, 0, 0, 3, 0]
    out.append(record("take3_one_colour", s, 2, "tests/test_take_reduced_colors.py:24-36"))
    # tests/test_afford_nobles_obs.py:31-43 -- all bonuses 4: exactly one noble even on a take action
    s = R.initial_state(seed=999)
    s.players[0].bonuses = [4, 4, 4, 4, 4]
    out.append(record("one_noble_per_turn", s, [i for i, v in enumerate(R.legal_moves(s)) if v][0],
                      "tests/test_afford_nobles_obs.py:31-43; engine/rules.py:132-147"))
    # tests/test_afford_nobles_obs.py:9-28 restated with a real card: discounts + gold substitution
    s = R.initial_state(seed=123)
    s.board[1][0] = C[2]  # cost white2 blue2 red1, colour black
    s.players[0].tokens = [1, 2, 0, 0, 0, 2]
    s.players[0].bonuses = [0, 0, 0, 0, 0]
    out.append(record("buy_with_gold_substitution", s, 15, "tests/test_afford_nobles_obs.py:9-28; engine/rules.py:101-122"))
    s = R.initial_state(seed=123)
    s.board[1][0] = C[2]
    s.players[0].tokens = [1, 2, 0, 0, 0, 1]
    out.append(record("buy_unaffordable_is_illegal", s, 15, "tests/test_gym_compat.py:111-124; envs/splendor_env.py:64-66"))
    s = R.initial_state(seed=123)
    s.board[3][2] = C[89]  # green7 red3
    s.players[0].tokens = [0, 0, 1, 0, 0, 3]
    s.players[0].bonuses = [0, 0, 4, 2, 9]
    out.append(record("buy_tier3_bonus_discount", s, 15 + 8 + 2, "engine/state.py:61-71"))
    # reserve visible / blind, gold exhausted, 3rd reservation
    s = R.initial_state(seed=11)
    s.bank[5] = 0
    out.append(record("reserve_visible_no_gold", s, 27 + 5, "engine/rules.py:226-240"))
    s = R.initial_state(seed=11)
    out.append(record("reserve_blind_tier3", s, 41, "engine/rules.py:241-249"))
    s = R.initial_state(seed=11)
    s.decks[3].clear()
    out.append(record("reserve_blind_empty_deck_illegal", s, 41, "engine/rules.py:83-86"))
    s = R.initial_state(seed=11)
    s.decks[2].clear()
    out.append(record("buy_visible_no_refill", s, 15 + 4, "engine/rules.py:125-129"))
    s = R.initial_state(seed=11)
    s.decks[2].clear()
    s.players[0].tokens = [3, 3, 3, 3, 3, 2]
    s.players[0].bonuses = [3, 3, 3, 3, 3]
    out.append(record("buy_visible_empty_deck_leaves_hole", s, 15 + 4, "engine/rules.py:125-129,216-225"))
    # buy reserved: list.pop(idx) shifts ([33,35,37] -> [35,37])
    s = R.initial_state(seed=21)
    s.players[0].reserved = [C[33], C[35], C[37]]
    s.players[0].revealed_reserved = [True, False, True]
    s.players[0].tokens = [2, 2, 2, 2, 2, 0]
    s.players[0].bonuses = [2, 2, 2, 2, 2]
    out.append(record("buy_reserved_slot0_shifts", s, 42, "engine/rules.py:250-255"))
    s = R.initial_state(seed=21)
    s.players[0].reserved = [C[33], C[35], C[37]]
    s.players[0].revealed_reserved = [True, False, True]
    s.players[0].tokens = [2, 2, 2, 2, 2, 0]
    s.players[0].bonuses = [2, 2, 2, 2, 2]
    out.append(record("buy_reserved_slot1_shifts", s, 43, "engine/rules.py:250-255"))
    # opponent's hidden reservation is 14 zeros in the observation (tests/test_reserved_card_observation.py:112-138)
    s = R.initial_state(seed=31)
    s = R.apply_action(s, 39)   # P0 reserves blind tier 1
    s = R.apply_action(s, 27)   # P1 reserves visible
    out.append(record("opponent_hidden_reserved_obs", s, 40, "tests/test_reserved_card_observation.py:112-138"))
    # end of game: P0 reaches 15 -> game_over but not terminal; P1's reply terminates
    s = R.initial_state(seed=41)
    s.players[0].prestige = 14
    s.players[0].bonuses = [7, 7, 7, 7, 7]
    s.board[1][0] = C[7]  # 1 point
    st_mid = record("p0_reaches_15_not_terminal", s, 15, "engine/rules.py:263-285")
    out.append(st_mid)
    s2 = pyref.row_to_state(st_mid["row_out"])
    out.append(record("p1_last_move_terminal_p0_wins", s2, 0, "envs/splendor_env.py:68-88"))
    # P1 reaches 15 first: immediate terminal, mover wins -> +1
    s = R.initial_state(seed=41)
    s = R.apply_action(s, 0)
    s.players[1].prestige = 14
    s.players[1].bonuses = [7, 7, 7, 7, 7]
    s.board[1][1] = C[7]
    out.append(record("p1_reaches_15_wins", s, 16, "engine/rules.py:282-285"))
    # tie-breaks: equal prestige, fewer cards wins; exact tie -> None
    for name, b0, b1, r0 in (("tiebreak_fewer_cards", [3, 3, 3, 3, 3], [2, 3, 3, 3, 3], 0), ("tiebreak_exact_tie", [3, 3, 3, 3, 3], [3, 3, 3, 3, 3], 0),
                             ("tiebreak_fewer_reserved", [3, 3, 3, 3, 3], [3, 3, 3, 3, 3], 1)):
        s = R.initial_state(seed=51)
        s = R.apply_action(s, 0)
        s.game_over = True
        s.players[0].prestige = 16
        s.players[1].prestige = 16
        s.players[0].bonuses = list(b0)
        s.players[1].bonuses = list(b1)
        if r0:
            s.players[0].reserved = [C[50]]
            s.players[0].revealed_reserved = [True]
        out.append(record(name, s, 1, "engine/rules.py:290-303"))
    # turn limit: move 198 -> draw by limit overrides a real win
    s = R.initial_state(seed=61)
    s.move_count = 197
    s.turn_count = 99
    s.to_play = 1
    s.players[1].prestige = 20
    out.append(record("turn_limit_overrides_win", s, 0, "engine/rules.py:274-279; envs/splendor_env.py:71-75"))
    s = R.initial_state(seed=61)
    s.move_count = 196
    s.turn_count = 99
    s.to_play = 0
    out.append(record("move_197_not_terminal", s, 0, "engine/rules.py:268-279"))
    # step after termination raises (tests/test_gym_compat.py:89-108); out-of-range action raises
    s = R.initial_state(seed=61)
    s.game_over = True
    out.append(record("step_after_terminal", s, 0, "envs/splendor_env.py:53-54"))
    s = R.initial_state(seed=61)
    out.append(record("action_out_of_range", s, 45, "envs/splendor_env.py:62-63"))
    s = R.initial_state(seed=61)
    out.append(record("action_negative", s, -1, "envs/splendor_env.py:62-63"))
    # take-2 needs 4 (tests/test_rules.py:29-34)
    s = R.initial_state(seed=71)
    s.bank[2] = 3
    out.append(record("take2_needs_four_illegal", s, 12, "tests/test_rules.py:29-34"))
    s = R.initial_state(seed=71)
    out.append(record("take2_ok", s, 12, "engine/rules.py:211-215"))
    # in-domain token return with k=1,2,3 (hand 10 + take 3 / take 2 / reserve with gold)
    s = R.initial_state(seed=81)
    s.players[0].tokens = [2, 2, 2, 2, 2, 0]
    s.bank[:] = [2, 2, 2, 2, 2, 5]
    out.append(record("return_k3", s, 4, "engine/rules.py:150-193"))
    s = R.initial_state(seed=81)
    s.players[0].tokens = [0, 0, 3, 3, 3, 1]
    s.bank[:] = [4, 4, 1, 1, 1, 4]
    out.append(record("return_k2_take2", s, 10, "engine/rules.py:150-193"))
    s = R.initial_state(seed=81)
    s.players[0].tokens = [5, 0, 0, 0, 0, 5]
    out.append(record("return_k1_reserve_gold", s, 27, "engine/rules.py:150-193,236-239"))
    dump("edge_cases.json", out)




# ----------------------------------------------------------------------------- bots.json
def gen_bots():
    """(obs, mask) states from reference games + the decision of each scripted opponent of scripts/eval_suite.py.
    Deterministic bots: the action.  basic_priority (np.random.choice): the support set over 64 RNG seeds."""
    from splendor_gym.scripts import eval_suite as ES

    out = []
    rng = np.random.RandomState(5)
    for g in range(24):
        env = env_from_state(R.initial_state(seed=1000 + g))
        obs = ENC.encode_observation(env.state)
        info = {"action_mask": np.array(R.legal_moves(env.state), dtype=np.int8)}
        v2 = ES.greedy_opponent_v2_factory(env)
        bots = [ES.greedy_opponent_v1, ES.basic_priority_opponent, v2]
        for t in range(300):
            mask = info["action_mask"]
            if t % 6 == g % 6:
                support = set()
                for sd in range(64):
                    np.random.seed(sd)
                    support.add(int(ES.basic_priority_opponent(obs, info)))
                out.append({"obs": obs.tolist(), "mask": mask.tolist(), "greedy_v1": int(ES.greedy_opponent_v1(obs, info)),
                            "greedy_v2": int(v2(obs, info)), "basic_support": sorted(support)})
            np.random.seed(int(rng.randint(1 << 30)))
            a = bots[(g + t) % 3](obs, info) if mask.any() else 0
            obs, r, term, trunc, info = env.step(int(a))
            if term:
                break
    dump("bots.json", out)


# ----------------------------------------------------------------------------- games_digest.json
DIGEST_SEED0, DIGEST_GAMES = 10_000, 10_000


def step_record(obs, mask, reward, term, bits) -> bytes:
    """One step of a game digest: observation as bytes (every entry < 256), mask, float32 reward, terminated, info bits."""
    o = np.asarray(obs)
    assert o.max() < 256 and o.min() >= 0
    return o.astype(np.uint8).tobytes() + np.asarray(mask, np.int8).tobytes() + struct.pack("<fBB", float(reward), int(bool(term)), bits)


def lcg_game_digest(seed: int):
    """The "lcg" policy of play() above; returns (moves, winner or -1, steps, sha256 of all step records)."""
    env = env_from_state(R.initial_state(seed=seed))
    x = (seed * 2654435761) % 2**32
    h = hashlib.sha256()
    t = 0
    while True:
        mask = R.legal_moves(env.state)
        legal = [i for i, v in enumerate(mask) if v]
        x = (1664525 * x + 1013904223) % 2**32
        a = legal[(x >> 16) % len(legal)] if legal else 0
        obs, r, term, trunc, info = env.step(a)
        h.update(step_record(obs, info["action_mask"], r, term, info_bits(info, env.state)))
        t += 1
        if term or t >= 400:
            break
    s = env.state
    return [s.move_count, -1 if s.winner_index is None else s.winner_index, t, h.hexdigest()[:16]]


def gen_games_digest():
    import multiprocessing as mp

    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
        rows = pool.map(lcg_game_digest, range(DIGEST_SEED0, DIGEST_SEED0 + DIGEST_GAMES), chunksize=50)
    dump("games_digest.json", {
        "policy": "lcg: x0 = seed * 2654435761 mod 2^32; every step x = 1664525 x + 1013904223 mod 2^32, action = legal[(x >> 16) % len(legal)] (0 if none)",
        "record": "per step: obs as uint8[297] | mask int8[45] | reward float32 LE | terminated u8 | info bits u8 "
                  "(1 illegal, 2 draw, 4 turn limit, 8 terminal, (winner + 1) << 4 with final_rewards); sha256 over all steps, first 16 hex digits",
        "seed0": DIGEST_SEED0, "columns": ["moves", "winner", "steps", "sha"], "games": rows,
        "env_steps": sum(r[2] for r in rows)})


def gen_games_digest_100k():
    """The same for 100,000 games (seeds 1,000,000 ...), stored compactly: one sha256 per CHUNK of 100 games over the games'
    (moves, winner, steps, sha) tuples, plus the totals.  ~10 min on 8 cores."""
    import multiprocessing as mp

    seed0, games, chunk = 1_000_000, 100_000, 100
    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
        rows = pool.map(lcg_game_digest, range(seed0, seed0 + games), chunksize=100)
    chunks = []
    for c in range(0, games, chunk):
        h = hashlib.sha256()
        for r in rows[c:c + chunk]:
            h.update(("%d,%d,%d,%s;" % tuple(r)).encode())
        chunks.append(h.hexdigest()[:16])
    dump("games_digest_100k.json", {
        "policy": "lcg (see games_digest.json)", "record": "see games_digest.json", "seed0": seed0, "games": games, "chunk": chunk,
        "chunk_digest": "sha256 over '%d,%d,%d,%s;' % (moves, winner, steps, game sha) of the chunk's games, first 16 hex digits",
        "chunks": chunks, "env_steps": sum(r[2] for r in rows), "max_steps": max(r[2] for r in rows),
        "winners": [sum(1 for r in rows if r[1] == w) for w in (0, 1, -1)]})


# ----------------------------------------------------------------------------- wrappers.json
def pick(obs, mask, mul: int, add: int) -> int:
    """Deterministic stand-in policy, a function of (obs, mask) only: the k-th legal action, k = (sum(obs) * mul + add) mod #legal."""
    legal = np.flatnonzero(np.asarray(mask))
    if len(legal) == 0:
        return 0
    return int(legal[(int(np.asarray(obs).sum()) * mul + add) % len(legal)])


def opponent_det(obs, info):
    return pick(obs, info["action_mask"], 7, 3)


def wrapper_game(kind: str, seed: int, late: bool):
    from splendor_gym.wrappers.dual_step_native import DualStepNativeWrapper
    from splendor_gym.wrappers.dual_step_selfplay import DualStepSelfPlayWrapper
    from splendor_gym.wrappers.selfplay import SelfPlayWrapper

    W = {"selfplay": SelfPlayWrapper, "dual_native": DualStepNativeWrapper, "dual_selfplay": DualStepSelfPlayWrapper}[kind]
    env = W(ENV.SplendorEnv(), opponent_policy=opponent_det, random_starts=False)
    obs, info = env.reset(seed=seed)
    base = env.env
    if late is True:  # a game that is already at move 190: the turn limit (engine/rules.py:272-277) ends it within a few turns
        base.state.move_count = 190
        base.state.turn_count = 96
        obs, info = ENC.encode_observation(base.state), {"action_mask": np.array(R.legal_moves(base.state), np.int8), "to_play": 0}
    if late == "stuck":  # the agent has no legal move: empty bank, three unaffordable reserved cards, nothing in hand
        st = base.state
        st.players[0].tokens = [0] * 6
        st.players[0].reserved = [st.decks[3].pop() for _ in range(3)]
        st.players[1].tokens = [t + b for t, b in zip(st.players[1].tokens, st.bank)]
        st.bank = [0] * 6
        obs, info = ENC.encode_observation(st), {"action_mask": np.array(R.legal_moves(st), np.int8), "to_play": 0}
        assert info["action_mask"].sum() == 0
    row0 = pyref.state_to_row(base.state).tolist()
    steps = []
    t = 0
    while True:
        a = pick(obs, info["action_mask"], 5, t)
        obs, r, term, trunc, info = env.step(a)
        rec = {"a": a, "r": float(r), "done": bool(term or trunc), "obs": hashlib.sha256(np.asarray(obs, np.int32).tobytes()).hexdigest()[:12],
               "mask": hashlib.sha256(np.asarray(info["action_mask"], np.int8).tobytes()).hexdigest()[:12],
               "row": hashlib.sha256(pyref.state_to_row(base.state).tobytes()).hexdigest()[:12]}
        for k in ("opponent_action", "opponent_reward", "game_ended_on", "phase", "turn_limit", "draw"):
            if k in info:
                rec[k] = info[k]
        if "final_rewards" in info:
            rec["final_rewards"] = [info["final_rewards"][0], info["final_rewards"][1]]
        steps.append(rec)
        t += 1
        if term or trunc or t >= 300:
            break
    out = {"wrapper": kind, "seed": seed, "late": late, "row0": row0, "steps": steps}
    if hasattr(env, "get_wrapper_stats"):
        out["stats"] = {k: v for k, v in env.get_wrapper_stats().items() if not isinstance(v, float)}
    return out


def gen_wrappers():
    out = []
    for kind in ("selfplay", "dual_native", "dual_selfplay"):
        for seed in range(300, 310):
            out.append(wrapper_game(kind, seed, False))
        for seed in range(320, 326):
            out.append(wrapper_game(kind, seed, True))
        out.append(wrapper_game(kind, 330, "stuck"))
        assert out[-1]["steps"][-1].get("draw") and len(out[-1]["steps"]) == 1
    limit = [g for g in out if g["steps"][-1].get("turn_limit")]
    assert len(limit) >= 3, "no turn-limit draw among the late games"
    # the quirk this fixture pins: a turn-limit draw on the opponent's move is +0.1 under SelfPlayWrapper (selfplay.py:55-57) and
    # -0.1 under the dual-step wrappers (final_rewards[0], dual_step_selfplay.py:138-152 / dual_step_native.py:159-161)
    assert any(g["wrapper"] == "selfplay" and g["steps"][-1]["r"] == 0.1 for g in limit)
    assert any(g["wrapper"] == "dual_selfplay" and g["steps"][-1]["r"] == -0.1 for g in limit)
    assert any(g["wrapper"] == "dual_native" and g["steps"][-1]["r"] == -0.1 for g in limit)
    dump("wrappers.json", out)


# ----------------------------------------------------------------------------- autoreset_stream.json
def gen_autoreset_stream():
    """A reference SplendorEnv that is reset(seed=s) once and then auto-reset four times the way the vector env does it
    (env.reset() without a seed: the next draw of the env's own PCG64 stream, envs/splendor_env.py:42-43,
    ppo_splendor.py:246-247).  Engine seeds drawn, start row of every episode, actions, per-step digests in the
    same-step auto-reset convention (terminal step: reward / terminated of the finished game, observation / mask / row of
    the new one, info bit 128)."""
    out = []
    for s in (0, 1, 7, 42, 2024, 31337, 99991, 123456789):
        env = ENV.SplendorEnv()
        twin = np.random.Generator(np.random.PCG64(np.random.SeedSequence(s)))  # gymnasium.utils.seeding.np_random(s)
        obs, info = env.reset(seed=s)
        engine_seeds = [int(twin.integers(0, 2**31 - 1))]
        starts = [pyref.state_to_row(env.state).tolist()]
        assert starts[0] == pyref.state_to_row(R.initial_state(seed=engine_seeds[0])).tolist()
        mask = info["action_mask"]
        x = (s * 2654435761 + 12345) % 2**32
        actions, digests = [], []
        while len(starts) < 5:
            legal = np.flatnonzero(mask)
            x = (1664525 * x + 1013904223) % 2**32
            a = int(legal[(x >> 16) % len(legal)]) if len(legal) else 0
            obs, r, term, trunc, info = env.step(a)
            bits = info_bits(info, env.state)
            mask = info["action_mask"]
            if term:
                obs, rinfo = env.reset()
                mask = rinfo["action_mask"]
                bits |= 128
                engine_seeds.append(int(twin.integers(0, 2**31 - 1)))
                starts.append(pyref.state_to_row(env.state).tolist())
                assert starts[-1] == pyref.state_to_row(R.initial_state(seed=engine_seeds[-1])).tolist()
            actions.append(a)
            digests.append(digest(obs, mask, pyref.state_to_row(env.state), r, term, bits))
        out.append({"seed": s, "engine_seeds": engine_seeds, "starts": starts, "actions": actions, "digests": digests})
    dump("autoreset_stream.json", out)


# ----------------------------------------------------------------------------- logger_strings.json
def gen_logger_strings():
    """scripts/game_logger.py:98-170 (decode_action for all 45 actions) and engine/state.py:61-71 (PlayerState.can_afford for
    every card on the board / in hand) on positions from a played game, incl. reduced take-3s and empty slots."""
    sys.path.insert(0, pyref.REFERENCE_ROOT)
    from splendor_gym.scripts.game_logger import SplendorGameLogger

    lg = SplendorGameLogger()
    out = []
    for seed, stops in ((5, (0, 9, 25, 48)), (11, (14, 33, 60))):
        state = R.initial_state(seed=seed)
        x = seed
        t = 0
        while True:
            if t in stops:
                s = state
                if t == 25:  # reduced take-3 (two colours left) and an empty board slot
                    import copy

                    s = copy.deepcopy(state)
                    s.bank[0] = s.bank[2] = s.bank[4] = 0
                    s.board[1][2] = None
                elif t == 33:
                    import copy

                    s = copy.deepcopy(state)
                    s.bank[:5] = [0, 0, 0, 2, 0]
                me = s.players[s.to_play]
                cards = [c for tier in (1, 2, 3) for c in s.board[tier] if c is not None] + list(me.reserved)
                out.append({"row": pyref.state_to_row(s).tolist(), "actions": [lg.decode_action(a, s) for a in range(46)],
                            "can_afford": [[c.id, bool(me.can_afford(c)[0]), list(me.can_afford(c)[1])] for c in cards]})
            if t >= max(stops) or R.is_terminal(state):
                break
            mask = R.legal_moves(state)
            legal = [i for i, v in enumerate(mask) if v]
            x = (1664525 * x + 1013904223) % 2**32
            state = R.apply_action(state, legal[(x >> 16) % len(legal)])
            t += 1
    assert len(out) == 7
    dump("logger_strings.json", out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["mt", "token_return", "initial", "env_seeding", "games", "edges", "bots", "games_digest", "wrappers", "autoreset_stream", "logger_strings"]
    for name in which:
        globals()["gen_" + name]()
