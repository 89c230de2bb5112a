"""Host-compiled copy of the lane-local CUDA rules (splendor_gym_b200/csrc/spl_core.cuh, built by
tests/emu) against the oracle and the golden vectors.  This checks the per-lane logic of the kernels
on a machine without a GPU; the GPU parity tests proper are tests/test_gpu_*.py."""
import hashlib
import struct

import numpy as np
import pytest

from conftest import load_golden
import spl_emu_host as emu


def test_ret_table_matches_cpython_golden():
    g = load_golden("token_return.json")
    t = emu.ret_table()
    assert hashlib.sha256(t.astype("<u8").tobytes()).hexdigest() == g["sha256"]


def test_mt_block_restatement(oracle):
    rng = np.random.RandomState(3)
    seeds = [0, 1, 2**32 - 1, 2**32, 2**37 + 99] + [int(x) for x in rng.randint(0, 2**62, size=20)]
    for seed in seeds:
        outs = oracle.mt_outputs(seed, 210)
        for blk in (0, 1, 5, 9):
            want = 0
            for j in range(21):
                want |= (int(outs[21 * blk + j]) >> 29) << (3 * j)
            assert emu.mt_block(seed, blk) == want, (seed, blk)


def _deal_from_row(row):
    """(tier decks incl. the dealt cards, visible nobles) of a flat state row right after initial_state."""
    row = list(row)
    decks = []
    for t, (off, full) in enumerate(((76, 40), (116, 30), (146, 20))):
        rest = row[off:off + full - 4]
        board = row[52 + 4 * t:56 + 4 * t]       # board[t][k] = deck.pop() for k = 0..3 (engine/state.py:190-191)
        decks.append(rest + board[::-1])
    return decks, row[67:70]


def test_batch_dealer_body_matches_the_reference_deals(oracle):
    """spl_mt_deal_stream (MT19937 in registers: pass 1 re-run next to pass 2, outputs streamed from two chain copies,
    one flat shuffle loop) against initial_state(seed) of the reference itself (tests/golden/initial_states.json) and
    against the oracle for 300 more seeds; the output budget hands a deal back instead of producing a wrong one."""
    def check(seed, row):
        ok, deck = emu.mt_deal_stream(seed)
        assert ok, seed
        decks, nobles = _deal_from_row(row)
        assert deck[:40].tolist() == decks[0] and deck[40:70].tolist() == decks[1] and deck[70:90].tolist() == decks[2], seed
        assert deck[90:93].tolist() == nobles, seed

    for g in load_golden("initial_states.json"):
        check(g["seed"], g["row"])
    for base in (0, 123456, 2147483646 - 299):
        v = oracle.OracleVec(300, seed_base=base)
        v.reset()
        rows = v.export_rows()
        for i in range(300):
            check((base + i) % 2147483647, rows[i])
    used = []
    for seed in range(40):
        lo = next(m for m in range(90, 228) if emu.mt_deal_stream(seed, m)[0])  # smallest budget that completes the deal
        assert not emu.mt_deal_stream(seed, lo - 1)[0] and emu.mt_deal_stream(seed, 227)[0]
        used.append(lo)
    assert 100 <= min(used) and max(used) <= 200, used  # a deal draws ~130 outputs (96 accepted + rejections)


def test_edge_cases_golden():
    for c in load_golden("edge_cases.json"):
        row_in = np.array(c["row_in"], np.int32)
        assert emu.roundtrip(row_in).tolist() == c["row_in"], c["name"]
        obs0, mask0 = emu.observe(row_in)
        game_over_terminal = c.get("raises") == "RuntimeError"
        if not game_over_terminal:
            assert mask0.tolist() == c["mask_in"], c["name"]
        row, obs, mask, r, term, info = emu.env_step(row_in, c["action"])
        if "raises" in c:
            assert info & 64 and row.tolist() == c["row_in"], c["name"]
            continue
        assert row.tolist() == c["row_out"], c["name"]
        assert obs.tolist() == c["obs"], c["name"]
        assert mask.tolist() == c["mask"], c["name"]
        assert r == pytest.approx(c["reward"]) and term == c["terminated"] and info == c["info"], c["name"]


@pytest.mark.parametrize("idx", range(0, 36, 3))
def test_golden_games(idx):
    from test_oracle_golden import digest

    g = load_golden("games.json")[idx]
    import oracle.oracle as O

    row = O.initial_row(g["seed"])
    for t, a in enumerate(g["actions"]):
        row, obs, mask, r, term, info = emu.env_step(row, a)
        assert digest(obs, mask, row, r, term, info) == g["digests"][t], (g["seed"], g["policy"], t)


def test_random_games_vs_oracle(oracle):
    rng = np.random.RandomState(11)
    steps = 0
    for g in range(150):
        row = oracle.initial_row(int(rng.randint(0, 2**31 - 1)))
        for t in range(400):
            obs0, mask0 = emu.observe(row)
            assert np.array_equal(mask0, oracle.legal_moves(row))
            assert np.array_equal(obs0, oracle.encode_observation(row))
            legal = np.flatnonzero(mask0)
            if len(legal) == 0:
                a = 0
            elif rng.rand() < 0.04:
                a = int(rng.randint(-2, 48))
            else:
                a = int(legal[rng.randint(len(legal))])
            want = oracle.env_step(row, a)
            got = emu.env_step(row, a)
            for x, y in zip(want[:3], got[:3]):
                assert np.array_equal(x, y), (g, t, a)
            assert want[3] == pytest.approx(got[3]) and want[4] == got[4] and want[5] == got[5], (g, t, a)
            row = want[0]
            steps += 1
            if want[4]:
                break
    assert steps > 8000


def test_rollout_chunking_partitions_the_steps():
    """The rollout kernel's work units: chunks tile [0, steps) exactly, in order, full-size first and a halving tail."""
    for steps in list(range(1, 70)) + [128, 150, 1000]:
        for chunk in (1, 2, 3, 8, 16, 64):
            ch = emu.chunks(steps, chunk)
            pos = 0
            for s, n in ch:
                assert s == pos and n >= 1
                pos += n
            assert pos == steps
            assert all(n <= 2 * chunk for _, n in ch)
            assert ch[-1][1] <= 3  # short last unit -> short tail of the launch
            assert len(ch) <= steps // chunk + 8


def test_row_validation_accepts_every_reference_row_and_rejects_what_would_be_truncated():
    """spl_row_valid (the admission rule of spl_import_state): every row the reference produced -- initial states, the hand-built
    edge cases of its own tests, the wrapper / auto-reset fixtures -- is inside the packed state's domain; rows whose counters
    would not fit the byte-wise arithmetic (>= 128), whose ids fall outside the card / noble tables or whose lists are too long
    are rejected instead of being masked to a byte."""
    from conftest import load_golden
    import spl_emu_host as E

    rows = [np.array(c["row_in"], np.int32) for c in load_golden("edge_cases.json") if "row_in" in c]
    rows += [np.array(g["row0"], np.int32) for g in load_golden("wrappers.json")]
    rows += [np.array(r, np.int32) for g in load_golden("autoreset_stream.json") for r in g["starts"]]
    rows += [np.array(p["row"], np.int32) for p in load_golden("logger_strings.json")]
    assert len(rows) > 100 and all(E.row_valid(r) for r in rows)
    base = rows[-1]
    for col, val in ((0, 128), (0, -1), (6, 200), (17, 128), (18, 4), (52, 90), (52, -2), (64, 41), (65, 31), (66, 21), (67, 10),
                     (70, 2), (71, 256), (74, 2), (76, 90)):
        bad = base.copy()
        bad[col] = val
        assert not E.row_valid(bad), (col, val)
    ok = base.copy()
    ok[0], ok[17], ok[71] = 127, 127, 255  # the largest admissible counters
    assert E.row_valid(ok)
