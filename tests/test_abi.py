"""The C-ABI library loads on a machine without a GPU and exports every function include/splendor_b200.h declares;
the ctypes mirrors of the ABI structs have the C layout; host-only entry points work (no compute is launched)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from splendor_gym_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "splendor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int|double)\s*\*?\s*(spl_[a-z_0-9]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(lib):
    from splendor_gym_b200 import _lib

    names = declared_functions()
    assert len(names) >= 18 and "spl_step" in names and "spl_rollout_random" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/splendor_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "splendor_gym_b200/_lib.py EXPORTS out of sync with the header"


def test_struct_layouts_match_header():
    from splendor_gym_b200 import _lib

    # struct spl_envs: 4 pointers, 2 int64, 2 uint64, 2 int32, 2 pointers, 2 int32 ; struct spl_step_io: 9 pointers, 2 uint64, 1 pointer, 2 int32
    assert C.sizeof(_lib.SplEnvs) == 4 * 8 + 4 * 8 + 8 + 2 * 8 + 8 and _lib.SplEnvs.spare.offset == 72
    assert _lib.SplEnvs.episode_seeds.offset == 80 and _lib.SplEnvs.episode_seed_count.offset == 88
    assert C.sizeof(_lib.SplStepIO) == 9 * 8 + 2 * 8 + 8 + 8 + 2 * 8 and _lib.SplStepIO.obs_f16.offset == 104
    assert _lib.SplEnvs.shuffle_mode.offset == 64 and _lib.SplStepIO.autoreset.offset == 96
    # struct spl_host_io: 9 pointers, 2 uint64, 2 int32
    assert C.sizeof(_lib.SplHostIO) == 9 * 8 + 2 * 8 + 8 and _lib.SplHostIO.autoreset.offset == 88
    text = open(os.path.join(ROOT, "include", "splendor_b200.h")).read()
    for name, val in (("SPL_NUM_ACTIONS", _lib.NUM_ACTIONS), ("SPL_OBS_DIM", _lib.OBS_DIM), ("SPL_ROW_LEN", _lib.ROW_LEN),
                      ("SPL_DECK_STRIDE", _lib.DECK_STRIDE), ("SPL_RET_TABLE_LEN", _lib.RET_TABLE_LEN),
                      ("SPL_OBS_F16_PITCH", _lib.OBS_F16_PITCH)):
        assert re.search(rf"#define {name} {val}\b", text), name
    for name, val in (("SPL_INFO_ILLEGAL", _lib.INFO_ILLEGAL), ("SPL_INFO_NOLEGAL_DRAW", _lib.INFO_NOLEGAL_DRAW),
                      ("SPL_INFO_TURN_LIMIT", _lib.INFO_TURN_LIMIT), ("SPL_INFO_TERMINATED", _lib.INFO_TERMINATED),
                      ("SPL_INFO_ERROR", _lib.INFO_ERROR), ("SPL_INFO_RESET", _lib.INFO_RESET)):
        assert re.search(rf"#define {name} {val}u\b", text), name


def test_host_only_entry_points(lib):
    assert lib.spl_version() >= 120
    assert lib.spl_error_string(0) == b"ok" and b"bad argument" in lib.spl_error_string(-1)
    assert lib.spl_launch_count() == 0  # nothing has been launched by loading the library
    t = np.zeros(8910, np.uint64)
    assert lib.spl_host_ret_table(t.ctypes.data) == 0
    assert hashlib.sha256(t.astype("<u8").tobytes()).hexdigest() == load_golden("token_return.json")["sha256"]


def test_argument_errors_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call (return codes, never exceptions / exits)."""
    from splendor_gym_b200 import _lib

    assert lib.spl_step(None, None, None) == -1
    envs = _lib.SplEnvs()
    assert lib.spl_reset(C.byref(envs), None, None, None, None, None) == -1
    assert lib.spl_random_action(None, 0, 0, 0, 0, None, None) == -1
    assert lib.spl_gae(None, None, None, None, 1, 1, 0.99, 0.95, None, None, None) == -1


@pytest.mark.parametrize("n,threads", [(1, 1), (33, 2), (1000, 4), (5001, 3)])
def test_host_expand_widens_compact_records(lib, n, threads):
    """spl_host_expand (host half of spl_host_step): compact records -> the reference-typed arrays, checked against
    a NumPy restatement of the record layout; ragged sizes exercise the unaligned head / tail of the SIMD loop."""
    from splendor_gym_b200 import _lib

    rng = np.random.default_rng(n)
    obs8 = rng.integers(0, 256, size=(n, 297), dtype=np.uint8)
    mbits = rng.integers(0, 1 << 45, size=n, dtype=np.uint64)
    code = rng.integers(0, 5, size=n, dtype=np.uint32)
    term = rng.integers(0, 2, size=n, dtype=np.uint32)
    info = rng.integers(0, 256, size=n, dtype=np.uint32)
    act = rng.integers(0, 45, size=n, dtype=np.uint32)
    side = np.zeros((n, 4), np.uint32)
    side[:, 0] = (mbits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    side[:, 1] = (mbits >> np.uint64(32)).astype(np.uint32) | (code << 16) | (term << 24)
    side[:, 2] = info | (act << 8)
    # offset the destination by one element so that it is NOT 32-byte aligned
    obs_store = np.full(n * 297 + 9, -7, np.int32)
    obs = obs_store[1:1 + n * 297].reshape(n, 297)
    mask = np.full((n, 45), 9, np.int8)
    guard = np.full(64, 9, np.int8)
    rew, te, inf, nxt = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
    obs8_out = np.zeros((n, 297), np.uint8)
    io = _lib.SplHostIO(obs=obs.ctypes.data, obs_u8=obs8_out.ctypes.data, mask=mask.ctypes.data, reward=rew.ctypes.data,
                        terminated=te.ctypes.data, info=inf.ctypes.data, next_action=nxt.ctypes.data)
    assert lib.spl_host_set_threads(threads) == threads
    assert lib.spl_host_expand(obs8.ctypes.data, side.ctypes.data, n, C.byref(io)) == 0
    assert np.array_equal(obs, obs8.astype(np.int32)) and obs_store[0] == -7 and np.all(obs_store[1 + n * 297:] == -7)
    assert np.array_equal(obs8_out, obs8)
    want_mask = ((mbits[:, None] >> np.arange(45, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8)
    assert np.array_equal(mask, want_mask) and np.all(guard == 9)
    assert np.array_equal(rew, np.array([0.0, 1.0, -1.0, -0.1, -0.01], np.float32)[code])
    assert np.array_equal(te, term.astype(np.uint8)) and np.array_equal(inf, info.astype(np.uint8))
    assert np.array_equal(nxt, act.astype(np.int32))
    assert lib.spl_host_expand(None, None, n, C.byref(io)) == -1


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under splendor_gym_b200/ may import, load or call it."""
    pkg = os.path.join(ROOT, "splendor_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, os.path.join(dirpath, f)


def test_no_cpu_fallback():
    import torch

    from splendor_gym_b200 import SplendorB200Error, SplendorVecEnv

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((SplendorB200Error, RuntimeError, AssertionError)):
        SplendorVecEnv(4, device="cpu")


def test_host_alloc_returns_aligned_zeroed_memory_even_without_a_device(lib):
    """spl_host_alloc: 2 MB-aligned, zeroed, writable host memory.  Without a CUDA device the pinning step fails and the block is
    handed out unpinned (spl_host_step would then widen everything on the host) -- allocation itself must not fail."""
    p = C.c_void_p()
    assert lib.spl_host_alloc(C.c_size_t(3 * 1188 * 64 + 5), C.byref(p)) == 0 and p.value
    assert p.value % (2 << 20) == 0
    buf = (C.c_uint8 * (3 * 1188 * 64 + 5)).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.uint8)
    assert not a.any()
    a[:] = 7
    assert int(a.sum()) == 7 * a.size
    del a, buf
    assert lib.spl_host_free(p) == 0
    assert lib.spl_host_alloc(C.c_size_t(0), C.byref(p)) == -1
