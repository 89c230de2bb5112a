"""Pins the C oracle (oracle/splendor_oracle.c) against fixtures produced by executing the
reference itself (oracle/gen_golden.py -> tests/golden/).  CPU only."""
import hashlib
import struct

import numpy as np
import pytest

from conftest import load_golden


def digest(obs, mask, row, reward, term, bits):
    h = hashlib.sha256()
    h.update(np.asarray(obs, np.int32).tobytes())
    h.update(np.asarray(mask, np.int8).tobytes())
    h.update(np.asarray(row, np.int32).tobytes())
    h.update(struct.pack("<fBB", float(reward), int(bool(term)), bits))
    return h.hexdigest()[:12]


def test_mt19937_matches_cpython(oracle):
    for rec in load_golden("mt19937.json"):
        outs = oracle.mt_outputs(int(rec["seed"]), 1000)
        assert outs[:8].tolist() == rec["first"]
        assert hashlib.sha256(struct.pack("<1000I", *outs.tolist())).hexdigest() == rec["sha_1000"]


def test_token_return_streams(oracle):
    g = load_golden("token_return.json")
    h = hashlib.sha256()
    samples = {}
    for turn in range(1, 100):
        for tp in range(2):
            for hand in range(11, 14):
                for bank in range(15):
                    seed = (turn * 1315423911) ^ (tp * 2654435761) ^ (hand * 97531) ^ (bank * 31337)
                    o = oracle.mt_outputs(seed, 21)
                    v = 0
                    for j in range(21):
                        v |= (int(o[j]) >> 29) << (3 * j)
                    h.update(struct.pack("<Q", v))
                    samples[f"{turn},{tp},{hand},{bank}"] = str(v)
    assert h.hexdigest() == g["sha256"]
    for k, v in g["samples"].items():
        assert samples[k] == v


def test_initial_states(oracle):
    for rec in load_golden("initial_states.json"):
        row = oracle.initial_row(rec["seed"])
        assert row.tolist() == rec["row"], rec["seed"]
        obs = oracle.encode_observation(row)
        assert hashlib.sha256(obs.tobytes()).hexdigest()[:16] == rec["obs_sha16"]
        assert oracle.legal_moves(row).tolist() == rec["mask"]


def test_survey_fingerprints(oracle):
    """values obtained by the surveyor running the reference."""
    want = {
        0: ([[24, 26, 2, 16], [54, 67, 56, 48], [86, 85, 73, 79]], [7, 9, 5], [32, 41, 89], "acccb83d866e5f0c"),
        42: ([[7, 1, 17, 15], [51, 67, 69, 59], [81, 75, 89, 87]], [4, 9, 1], [14, 48, 76], "1e1673d87f5d6bc3"),
        123456789: ([[28, 34, 25, 19], [43, 63, 54, 40], [89, 86, 82, 76]], [3, 7, 2], [29, 48, 79], "a17c0ca2aaf0fcfc"),
    }
    for seed, (board, nobles, tops, sha) in want.items():
        row = oracle.initial_row(seed)
        assert row[52:64].reshape(3, 4).tolist() == board
        assert row[67:70].tolist() == nobles
        off = [76, 116, 146]
        assert [int(row[off[t] + row[64 + t] - 1]) for t in range(3)] == tops
        assert hashlib.sha256(oracle.encode_observation(row).tobytes()).hexdigest()[:16] == sha


def test_env_seeding_expression():
    """gymnasium's np_random(seed) == Generator(PCG64(SeedSequence(seed))); engine seed = integers(0, 2**31-1)
    (envs/splendor_env.py:42-43).  Pinned values from the reference run with the stand-in."""
    for rec in load_golden("env_seeding.json"):
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(rec["seed"])))
        assert int(g.integers(0, 2**31 - 1)) == rec["engine_seed"]
        assert int(g.integers(0, 2**31 - 1)) == rec["engine_seed_2nd_reset"]
    want = {0: 1826701614, 42: 191664963, 123: 33158374}
    got = {r["seed"]: r["engine_seed"] for r in load_golden("env_seeding.json")}
    for k, v in want.items():
        assert got[k] == v


@pytest.mark.parametrize("idx", range(36))
def test_games(oracle, idx):
    games = load_golden("games.json")
    assert len(games) == 36
    g = games[idx]
    row = oracle.initial_row(g["seed"])
    for t, a in enumerate(g["actions"]):
        row, obs, mask, r, term, info = oracle.env_step(row, a)
        assert digest(obs, mask, row, r, term, info) == g["digests"][t], (g["seed"], g["policy"], t)
        if str(t) in g["full"]:
            f = g["full"][str(t)]
            assert obs.tolist() == f["obs"] and mask.tolist() == f["mask"] and row.tolist() == f["row"]
            assert r == pytest.approx(f["reward"]) and term == f["terminated"] and info == f["info"]
    assert int(row[72]) == g["moves"]
    assert (None if row[74] < 0 else int(row[74])) == g["winner"]
    assert [int(row[17]), int(row[40])] == g["prestige"]


def test_edge_cases(oracle):
    cases = load_golden("edge_cases.json")
    assert len(cases) >= 35
    for c in cases:
        row_in = np.array(c["row_in"], np.int32)
        assert oracle.legal_moves(row_in).tolist() == c["mask_in"], c["name"]
        row, obs, mask, r, term, info = oracle.env_step(row_in, c["action"])
        if "raises" in c:
            assert info & oracle.INFO_ERROR, c["name"]
            assert row.tolist() == c["row_in"], c["name"]  # state untouched
            assert bool(info & oracle.INFO_TERMINATED) == (c["raises"] == "RuntimeError")
            continue
        assert row.tolist() == c["row_out"], c["name"]
        assert obs.tolist() == c["obs"], c["name"]
        assert mask.tolist() == c["mask"], c["name"]
        assert r == pytest.approx(c["reward"]), c["name"]
        assert term == c["terminated"], c["name"]
        assert info == c["info"], c["name"]
        if c["final_rewards"] is not None:
            w = ((info >> 4) & 3) - 1
            lim = bool(info & 4)
            fr = [(-0.1 if lim else 0.0) if w < 0 else (1.0 if w == p else -1.0) for p in range(2)]
            assert fr == pytest.approx(c["final_rewards"]), c["name"]


def test_vec_lockstep_autoreset_matches_single(oracle):
    """OracleVec (batched, same-step auto-reset) == repeated single-state env_step + re-seeded reset."""
    n, T = 16, 220
    v = oracle.OracleVec(n, seed_base=5, env_offset=100)
    v.reset()
    rows = [oracle.initial_row(oracle.engine_seed(5, 100 + i, 0)) for i in range(n)]
    ep = [0] * n
    assert np.array_equal(v.export_rows(), np.stack(rows))
    for t in range(T):
        a = v.random_actions(key=0xB200, t=t)
        obs, rew, term, info, mask = v.step(a, autoreset=True)
        for i in range(n):
            row, o, m, r, te, inf = oracle.env_step(rows[i], int(a[i]))
            assert r == pytest.approx(float(rew[i])) and te == bool(term[i])
            if te:
                ep[i] += 1
                row = oracle.initial_row(oracle.engine_seed(5, 100 + i, ep[i]))
                o, m = oracle.encode_observation(row), oracle.legal_moves(row)
                inf |= oracle.INFO_RESET
            assert inf == int(info[i])
            assert np.array_equal(o, obs[i]) and np.array_equal(m, mask[i])
            rows[i] = row
    assert v.episodes().tolist() == ep and sum(ep) > 0
    assert v.stats()[0] == sum(ep)


def _oracle_digest_games(oracle, seeds, T):
    import digest_util as D

    n = len(seeds)
    v = oracle.OracleVec(n)
    _, mask = v.reset(seeds=seeds)
    x = D.lcg_seed(seeds)
    steps = np.zeros(n, np.int64)
    active = np.ones(n, bool)
    recs = np.zeros((T, n, D.REC), np.uint8)
    for t in range(T):
        x = D.lcg_next(x)
        a = D.lcg_pick(x, mask)
        obs, rew, term, info, mask = v.step(a, active=active.astype(np.uint8), autoreset=False)
        recs[t] = D.step_records(obs, mask, rew, term, info)
        steps += active
        active &= term == 0
    assert not active.any()
    rows = v.export_rows()
    return rows[:, 72].copy(), rows[:, 74].copy(), steps, D.game_shas(recs, steps)


def test_ten_thousand_reference_games_by_digest(oracle):
    """tests/golden/games_digest.json: 10,000 games (773,764 env-steps) played by the UNMODIFIED reference under the LCG
    policy, one sha256 per game over every step's observation, mask, reward, terminated flag and info bits.  The oracle
    replays them all in lock-step and must reproduce every digest, move count and winner."""
    G = load_golden("games_digest.json")
    games = np.array([g[:3] for g in G["games"]], np.int64)
    n = len(games)
    moves, winner, steps, shas = _oracle_digest_games(oracle, G["seed0"] + np.arange(n, dtype=np.uint64), int(games[:, 2].max()))
    assert np.array_equal(steps, games[:, 2]) and np.array_equal(moves, games[:, 0]) and np.array_equal(winner, games[:, 1])
    bad = [i for i in range(n) if shas[i] != G["games"][i][3]]
    assert not bad, f"{len(bad)} of {n} game digests differ, first: seed {G['seed0'] + bad[0]}"


def test_hundred_thousand_reference_games_by_chunk_digest(oracle):
    """tests/golden/games_digest_100k.json: 100,000 further reference games (7.7e6 env-steps), stored as one digest per 100
    games.  Replayed by the oracle in slices of 20,000 envs."""
    import digest_util as D

    G = load_golden("games_digest_100k.json")
    got, total = [], 0
    for lo in range(0, G["games"], 20000):
        seeds = G["seed0"] + np.arange(lo, min(G["games"], lo + 20000), dtype=np.uint64)
        moves, winner, steps, shas = _oracle_digest_games(oracle, seeds, G["max_steps"])
        got += D.chunk_digests(moves, winner, steps, shas, G["chunk"])
        total += int(steps.sum())
    bad = [i for i, (a, b) in enumerate(zip(got, G["chunks"])) if a != b]
    assert not bad and len(got) == len(G["chunks"]), f"{len(bad)} of {len(got)} chunk digests differ, first chunk {bad[:1]}"
    assert total == G["env_steps"]


def test_logger_strings_and_can_afford_match_the_reference():
    """tests/golden/logger_strings.json: decode_action of the reference's scripts/game_logger.py for all 45 actions (+1 out of
    range) and PlayerState.can_afford for every visible / reserved card, on seven positions incl. reduced take-3s and empty
    slots.  The host mirror (no device needed) reproduces every string and every (affordable, cost_remaining) pair."""
    from splendor_gym_b200.engine.state import CARDS, row_to_state
    from splendor_gym_b200.trajectory import describe_action

    for pos in load_golden("logger_strings.json"):
        s = row_to_state(np.array(pos["row"], np.int32))
        assert [describe_action(a, s) for a in range(46)] == pos["actions"]
        assert describe_action(7, pos["row"]) == pos["actions"][7]
        me = s.players[s.to_play]
        for cid, ok, owed in pos["can_afford"]:
            got_ok, got_owed = me.can_afford(CARDS[cid])
            assert (bool(got_ok), list(got_owed)) == (ok, owed), cid
