"""ctypes binding of the CPU parity oracle (oracle/splendor_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

NUM_ACTIONS = 45
OBS_DIM = 297
ROW_LEN = 166

INFO_ILLEGAL = 1
INFO_NOLEGAL_DRAW = 2
INFO_TURN_LIMIT = 4
INFO_TERMINATED = 8
INFO_WINNER_SHIFT = 4
INFO_ERROR = 64
INFO_RESET = 128

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "splendor_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _p(a, ct):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ct))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_vec_create.restype = C.c_void_p
        L.orc_vec_create.argtypes = [C.c_int64, C.c_uint64, C.c_uint64]
        L.orc_vec_destroy.argtypes = [C.c_void_p]
        L.orc_engine_seed.restype = C.c_uint64
        L.orc_engine_seed.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_vec_rollout_random.restype = C.c_int64
        L.orc_vec_rollout_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64]
        L.orc_vec_random_actions.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleVec:
    """N independent reference-semantics environments stepped in lock-step on the host."""

    def __init__(self, n: int, seed_base: int = 0, env_offset: int = 0):
        self.n = int(n)
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_vec_create(self.n, seed_base, env_offset))
        self.obs = np.zeros((self.n, OBS_DIM), np.int32)
        self.mask = np.zeros((self.n, NUM_ACTIONS), np.int8)
        self.reward = np.zeros(self.n, np.float32)
        self.terminated = np.zeros(self.n, np.uint8)
        self.info = np.zeros(self.n, np.uint8)

    def __del__(self):
        try:
            self.L.orc_vec_destroy(self.h)
        except Exception:
            pass

    def reset(self, seeds=None, reset_mask=None):
        sd = None if seeds is None else np.ascontiguousarray(seeds, np.uint64)
        rm = None if reset_mask is None else np.ascontiguousarray(reset_mask, np.uint8)
        self.L.orc_vec_reset(self.h, _p(sd, C.c_uint64), _p(rm, C.c_uint8), _p(self.obs, C.c_int32), _p(self.mask, C.c_int8))
        return self.obs, self.mask

    def step(self, actions, active=None, autoreset=False):
        a = np.ascontiguousarray(actions, np.int32)
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        self.L.orc_vec_step(
            self.h, _p(a, C.c_int32), _p(act, C.c_uint8), int(bool(autoreset)),
            _p(self.obs, C.c_int32), _p(self.mask, C.c_int8), _p(self.reward, C.c_float),
            _p(self.terminated, C.c_uint8), _p(self.info, C.c_uint8),
        )
        return self.obs, self.reward, self.terminated, self.info, self.mask

    def observe(self):
        self.L.orc_vec_observe(self.h, _p(self.obs, C.c_int32), _p(self.mask, C.c_int8))
        return self.obs, self.mask

    def export_rows(self):
        rows = np.zeros((self.n, ROW_LEN), np.int32)
        self.L.orc_vec_export(self.h, _p(rows, C.c_int32))
        return rows

    def import_rows(self, rows, which=None):
        r = np.ascontiguousarray(rows, np.int32)
        w = None if which is None else np.ascontiguousarray(which, np.uint8)
        self.L.orc_vec_import(self.h, _p(r, C.c_int32), _p(w, C.c_uint8))

    def episodes(self):
        e = np.zeros(self.n, np.uint32)
        self.L.orc_vec_episodes(self.h, _p(e, C.c_uint32))
        return e

    def stats(self):
        s = np.zeros(8, np.int64)
        self.L.orc_vec_stats(self.h, _p(s, C.c_int64))
        return s

    def random_actions(self, key: int, t: int):
        a = np.zeros(self.n, np.int32)
        self.L.orc_vec_random_actions(self.h, key, t, a.ctypes.data_as(C.c_void_p))
        return a

    def rollout_random(self, key: int, t0: int, steps: int) -> int:
        return int(self.L.orc_vec_rollout_random(self.h, key, t0, steps))


# ---- single-state helpers (known-answer tests) ----
def initial_row(seed: int) -> np.ndarray:
    row = np.zeros(ROW_LEN, np.int32)
    lib().orc_initial_row(C.c_uint64(seed), _p(row, C.c_int32))
    return row


def legal_moves(row) -> np.ndarray:
    r = np.ascontiguousarray(row, np.int32)
    m = np.zeros(NUM_ACTIONS, np.int8)
    lib().orc_row_legal_moves(_p(r, C.c_int32), _p(m, C.c_int8))
    return m


def encode_observation(row) -> np.ndarray:
    r = np.ascontiguousarray(row, np.int32)
    o = np.zeros(OBS_DIM, np.int32)
    lib().orc_row_encode(_p(r, C.c_int32), _p(o, C.c_int32))
    return o


def apply_action(row, action: int) -> np.ndarray:
    r = np.ascontiguousarray(row, np.int32)
    out = np.zeros(ROW_LEN, np.int32)
    rc = lib().orc_row_apply(_p(r, C.c_int32), int(action), _p(out, C.c_int32))
    if rc != 0:
        raise ValueError("Invalid action index")
    return out


def env_step(row, action: int):
    """SplendorEnv.step on a flat row -> (row', obs, mask, reward, terminated, info)."""
    r = np.ascontiguousarray(row, np.int32)
    out = np.zeros(ROW_LEN, np.int32)
    obs = np.zeros(OBS_DIM, np.int32)
    mask = np.zeros(NUM_ACTIONS, np.int8)
    rew = C.c_float()
    term = C.c_uint8()
    info = C.c_uint8()
    lib().orc_row_env_step(_p(r, C.c_int32), int(action), _p(out, C.c_int32), _p(obs, C.c_int32), _p(mask, C.c_int8),
                           C.byref(rew), C.byref(term), C.byref(info))
    return out, obs, mask, float(rew.value), bool(term.value), int(info.value)


def mt_outputs(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.uint32)
    lib().orc_mt_outputs(C.c_uint64(seed), n, _p(out, C.c_uint32))
    return out


def philox(ctr, k0: int, k1: int) -> np.ndarray:
    c = np.ascontiguousarray(ctr, np.uint32).copy()
    lib().orc_philox(_p(c, C.c_uint32), C.c_uint32(k0), C.c_uint32(k1))
    return c


def engine_seed(seed_base: int, global_env: int, episode: int) -> int:
    return int(lib().orc_engine_seed(seed_base, global_env, episode))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))
