"""Live cross-check of the C oracle against the unmodified Python reference.  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import random

import numpy as np
import pytest

from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="reference tree not present")


def test_mt_vs_cpython_random(oracle):
    rng = random.Random(1234)
    for _ in range(40):
        seed = rng.getrandbits(rng.choice([1, 8, 31, 32, 33, 40, 63]))
        r = random.Random(seed)
        assert [r.getrandbits(32) for _ in range(700)] == oracle.mt_outputs(seed, 700).tolist()


def _check_games(span):
    """Games [lo, hi): every output of every step of the live reference against the oracle.  Returns env-steps checked."""
    from oracle import oracle as O

    lo, hi = span
    ns = pyref.load()
    steps = 0
    for g in range(lo, hi):
        rng = np.random.RandomState(g)
        seed = int(rng.randint(0, 2**31 - 1))
        env = ns.env.SplendorEnv()
        env.reset(seed=g)
        env.state = ns.rules.initial_state(seed=seed)
        row = O.initial_row(seed)
        assert np.array_equal(pyref.state_to_row(env.state), row)
        for t in range(500):
            mask = np.array(ns.rules.legal_moves(env.state), dtype=np.int8)
            assert np.array_equal(mask, O.legal_moves(row))
            legal = np.flatnonzero(mask)
            if len(legal) == 0:
                a = 0
            elif rng.rand() < 0.03:
                a = int(rng.randint(0, 45))
            else:
                a = int(legal[rng.randint(len(legal))])
            obs, r, term, trunc, info = env.step(a)
            row, obs2, mask2, r2, term2, info2 = O.env_step(row, a)
            assert np.array_equal(obs, obs2) and np.array_equal(info["action_mask"], mask2), (g, t)
            assert abs(r - r2) < 1e-6 and term == term2, (g, t)
            assert np.array_equal(pyref.state_to_row(env.state), row), (g, t)
            assert bool(info.get("illegal_action", False)) == bool(info2 & 1)
            assert bool(info.get("draw", False)) == bool(info2 & 2)
            assert bool(info.get("turn_limit", False)) == bool(info2 & 4)
            steps += 1
            if term:
                break
    return steps


def test_random_games_every_output(oracle):
    """2,000 random games by default (~1.5e5 env-steps, 3 % arbitrary -- often illegal -- actions), all cores;
    SPLENDOR_PYREF_GAMES=100000 for the long run (its log: profiles/r02_oracle_vs_pyref_1e5.log)."""
    import multiprocessing as mp
    import os

    games = int(os.environ.get("SPLENDOR_PYREF_GAMES", "2000"))
    procs = max(1, len(os.sched_getaffinity(0)))
    chunk = max(1, min(50, games // procs))
    spans = [(lo, min(games, lo + chunk)) for lo in range(0, games, chunk)]
    with mp.get_context("fork").Pool(procs) as pool:
        steps = sum(pool.map(_check_games, spans))
    print(f"oracle == live reference on {games} games, {steps} env-steps")
    assert steps > 50 * games


def test_row_roundtrip_through_reference_dataclass(oracle):
    ns = pyref.load()
    s = ns.rules.initial_state(seed=77)
    for a in (39, 27, 0, 1, 41, 30):
        if ns.rules.legal_moves(s)[a]:
            s = ns.rules.apply_action(s, a)
    row = pyref.state_to_row(s)
    s2 = pyref.row_to_state(row)
    assert np.array_equal(pyref.state_to_row(s2), row)
    assert np.array_equal(ns.encode.encode_observation(s2), oracle.encode_observation(row))
