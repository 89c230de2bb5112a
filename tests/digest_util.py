"""Helpers for tests/golden/games_digest.json (10,000 reference games, one sha256 per game; oracle/gen_golden.py:gen_games_digest)."""
import hashlib

import numpy as np

REC = 297 + 45 + 4 + 1 + 1  # obs as bytes | mask | float32 reward | terminated | info bits


def lcg_seed(seeds):
    return (np.asarray(seeds, np.uint64) * np.uint64(2654435761)) % np.uint64(2**32)


def lcg_next(x):
    return (np.uint64(1664525) * x + np.uint64(1013904223)) % np.uint64(2**32)


def lcg_pick(x, mask):
    """action = legal[(x >> 16) % len(legal)], 0 when there is no legal action (vectorised over envs)."""
    cnt = mask.sum(1).astype(np.int64)
    k = ((x >> np.uint64(16)).astype(np.int64)) % np.maximum(cnt, 1)
    a = (np.cumsum(mask, 1) > k[:, None]).argmax(1).astype(np.int32)
    a[cnt == 0] = 0
    return a


def step_records(obs, mask, reward, term, info):
    """[n, REC] uint8 records of one lock-step."""
    n = obs.shape[0]
    assert obs.min() >= 0 and obs.max() < 256
    rec = np.empty((n, REC), np.uint8)
    rec[:, :297] = obs
    rec[:, 297:342] = mask.view(np.uint8)
    rec[:, 342:346] = np.ascontiguousarray(reward, np.float32).view(np.uint8).reshape(n, 4)
    rec[:, 346] = term
    rec[:, 347] = info
    return rec


def game_shas(records, steps):
    """records [T, n, REC] uint8, steps [n] -> first 16 hex digits of sha256 over each game's first steps[i] records."""
    by_game = np.ascontiguousarray(records.transpose(1, 0, 2))
    return [hashlib.sha256(by_game[i, : steps[i]].tobytes()).hexdigest()[:16] for i in range(by_game.shape[0])]


def chunk_digests(moves, winner, steps, shas, chunk):
    """games_digest_100k.json: sha256 over '%d,%d,%d,%s;' % (moves, winner, steps, game sha) of every `chunk` games."""
    out = []
    for c in range(0, len(shas), chunk):
        h = hashlib.sha256()
        for i in range(c, min(len(shas), c + chunk)):
            h.update(("%d,%d,%d,%s;" % (int(moves[i]), int(winner[i]), int(steps[i]), shas[i])).encode())
        out.append(h.hexdigest()[:16])
    return out
