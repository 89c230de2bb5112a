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


def test_random_games_every_output(oracle):
    ns = pyref.load()
    rng = np.random.RandomState(0)
    steps = 0
    for g in range(60):
        seed = int(rng.randint(0, 2**31 - 1))
        env = ns.env.SplendorEnv()
        env.reset(seed=g)
        env.state = ns.rules.initial_state(seed=seed)
        row = oracle.initial_row(seed)
        assert np.array_equal(pyref.state_to_row(env.state), row)
        for t in range(500):
            mask = np.array(ns.rules.legal_moves(env.state), dtype=np.int8)
            assert np.array_equal(mask, oracle.legal_moves(row))
            legal = np.flatnonzero(mask)
            if len(legal) == 0:
                a = 0
            elif rng.rand() < 0.03:
                a = int(rng.randint(0, 45))
            else:
                a = int(legal[rng.randint(len(legal))])
            obs, r, term, trunc, info = env.step(a)
            row, obs2, mask2, r2, term2, info2 = oracle.env_step(row, a)
            assert np.array_equal(obs, obs2) and np.array_equal(info["action_mask"], mask2)
            assert r == pytest.approx(r2) and term == term2
            assert np.array_equal(pyref.state_to_row(env.state), row)
            assert bool(info.get("illegal_action", False)) == bool(info2 & 1)
            assert bool(info.get("draw", False)) == bool(info2 & 2)
            assert bool(info.get("turn_limit", False)) == bool(info2 & 4)
            steps += 1
            if term:
                break
    assert steps > 3000


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
