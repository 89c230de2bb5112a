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
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s*\*?\s*(spl_[a-z_0-9]+)\s*\(", text, flags=re.M)
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

    # struct spl_envs: 4 pointers, 2 int64, 2 uint64, 2 int32 ; struct spl_step_io: 9 pointers, 2 uint64, 1 pointer, 2 int32
    assert C.sizeof(_lib.SplEnvs) == 4 * 8 + 4 * 8 + 8
    assert C.sizeof(_lib.SplStepIO) == 9 * 8 + 2 * 8 + 8 + 8
    assert _lib.SplEnvs.shuffle_mode.offset == 64 and _lib.SplStepIO.autoreset.offset == 96
    text = open(os.path.join(ROOT, "include", "splendor_b200.h")).read()
    for name, val in (("SPL_NUM_ACTIONS", _lib.NUM_ACTIONS), ("SPL_OBS_DIM", _lib.OBS_DIM), ("SPL_ROW_LEN", _lib.ROW_LEN),
                      ("SPL_DECK_STRIDE", _lib.DECK_STRIDE), ("SPL_RET_TABLE_LEN", _lib.RET_TABLE_LEN)):
        assert re.search(rf"#define {name} {val}\b", text), name
    for name, val in (("SPL_INFO_ILLEGAL", _lib.INFO_ILLEGAL), ("SPL_INFO_NOLEGAL_DRAW", _lib.INFO_NOLEGAL_DRAW),
                      ("SPL_INFO_TURN_LIMIT", _lib.INFO_TURN_LIMIT), ("SPL_INFO_TERMINATED", _lib.INFO_TERMINATED),
                      ("SPL_INFO_ERROR", _lib.INFO_ERROR), ("SPL_INFO_RESET", _lib.INFO_RESET)):
        assert re.search(rf"#define {name} {val}u\b", text), name


def test_host_only_entry_points(lib):
    assert lib.spl_version() >= 100
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
