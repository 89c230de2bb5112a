"""ctypes loader for the test-only host build of the lane-local rules (tests/emu/spl_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libspl_emu.so")
ROOT = os.path.dirname(os.path.dirname(HERE))
_lib = None


def build():
    srcs = [os.path.join(HERE, "spl_emu.cpp"), os.path.join(ROOT, "splendor_gym_b200", "csrc", "spl_core.cuh"),
            os.path.join(ROOT, "splendor_gym_b200", "csrc", "spl_tables_host.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in srcs):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, srcs[0]])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emu_mt_block.restype = C.c_uint64
        _lib.emu_mt_block.argtypes = [C.c_uint64, C.c_uint32]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def env_step(row, action):
    r = np.ascontiguousarray(row, np.int32)
    out = np.zeros(166, np.int32)
    obs = np.zeros(297, np.int32)
    mask = np.zeros(45, np.int8)
    rew, term, info = C.c_float(), C.c_uint8(), C.c_uint8()
    lib().emu_env_step(_p(r, C.c_int32), int(action), _p(out, C.c_int32), _p(obs, C.c_int32), _p(mask, C.c_int8),
                       C.byref(rew), C.byref(term), C.byref(info))
    return out, obs, mask, float(rew.value), bool(term.value), int(info.value)


def observe(row):
    r = np.ascontiguousarray(row, np.int32)
    obs = np.zeros(297, np.int32)
    mask = np.zeros(45, np.int8)
    lib().emu_observe(_p(r, C.c_int32), _p(obs, C.c_int32), _p(mask, C.c_int8))
    return obs, mask


def mt_deal_stream(key, max_outputs=227):
    """(ok, deck[100]) of the batch dealer's per-thread body (spl_mt_deal_stream) for a one-word seed."""
    deck = np.zeros(100, np.uint8)
    ok = lib().emu_mt_deal_stream(C.c_uint32(int(key)), _p(deck, C.c_uint8), C.c_uint32(int(max_outputs)))
    return bool(ok), deck


def roundtrip(row):
    r = np.ascontiguousarray(row, np.int32)
    out = np.zeros(166, np.int32)
    lib().emu_roundtrip(_p(r, C.c_int32), _p(out, C.c_int32))
    return out


def ret_table():
    t = np.zeros(8910, np.uint64)
    lib().emu_ret_table(_p(t, C.c_uint64))
    return t


def mt_block(seed, blk):
    return int(lib().emu_mt_block(seed, blk))


def chunks(steps, chunk):
    """All (start, len) work-unit chunks of a `steps`-step rollout (spl_chunk_bounds)."""
    L = lib()
    out, c = [], 0
    s, n = C.c_int(), C.c_int()
    while L.emu_chunk_bounds(c, steps, chunk, C.byref(s), C.byref(n)):
        out.append((s.value, n.value))
        c += 1
    assert L.emu_num_chunks(steps, chunk) == c
    return out


def row_valid(row) -> bool:
    r = np.ascontiguousarray(row, np.int32)
    return bool(lib().emu_row_valid(_p(r, C.c_int32)))
