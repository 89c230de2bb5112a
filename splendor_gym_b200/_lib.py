"""ctypes binding of the C ABI declared in include/splendor_b200.h.

There is no CPU fallback: if libsplendor_b200.so is missing the import of the product path fails with
an explicit error (build it with ``python -m splendor_gym_b200.build`` / ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPL_LIB") or os.path.join(HERE, "libsplendor_b200.so")

NUM_ACTIONS = 45
OBS_DIM = 297
OBS_F16_PITCH = 304
ROW_LEN = 166
STATE_PLANES = 4
DECK_STRIDE = 96
RET_TABLE_LEN = 8910
MAX_SPARE_SLOTS = 16
IO_ASYNC_REFILL = 1

INFO_ILLEGAL = 1
INFO_NOLEGAL_DRAW = 2
INFO_TURN_LIMIT = 4
INFO_TERMINATED = 8
INFO_WINNER_SHIFT = 4
INFO_WINNER_MASK = 0x30
INFO_ERROR = 64
INFO_RESET = 128

SHUFFLE_MT19937 = 0
SHUFFLE_PHILOX = 1

STAT_NAMES = ("episodes", "p0_wins", "p1_wins", "tie_draws", "limit_draws", "nolegal_draws", "sum_moves", "sum_winner_prestige")

EXPORTS = (
    "spl_init", "spl_reset", "spl_step", "spl_observe", "spl_random_action", "spl_export_state", "spl_import_state",
    "spl_dual_combine", "spl_error_string", "spl_version", "spl_host_ret_table", "spl_launch_count",
    "spl_timing_enable", "spl_timing_read", "spl_rollout_random", "spl_scripted_action", "spl_masked_sample", "spl_gae",
    "spl_host_create", "spl_host_destroy", "spl_host_step", "spl_host_observe", "spl_host_set_threads", "spl_host_expand", "spl_rollout_plan", "spl_observe_policy", "spl_masked_sample_f16",
    "spl_refill_spares", "spl_host_set_pinning", "spl_host_alloc", "spl_host_free", "spl_host_get_stats", "spl_host_store_rate", "spl_load_deals",
)

BOT_RANDOM, BOT_GREEDY_V1, BOT_BASIC_PRIORITY, BOT_GREEDY_V2 = 0, 1, 2, 3


class SplEnvs(C.Structure):
    """struct spl_envs (include/splendor_b200.h)."""

    _fields_ = [
        ("state", C.c_void_p), ("decks", C.c_void_p), ("episode", C.c_void_p), ("scratch", C.c_void_p),
        ("stride", C.c_int64), ("n", C.c_int64), ("env_offset", C.c_uint64), ("seed_base", C.c_uint64),
        ("shuffle_mode", C.c_int32), ("spare_slots", C.c_int32), ("spare", C.c_void_p),
        ("episode_seeds", C.c_void_p), ("episode_seed_count", C.c_int32), ("reserved_", C.c_int32),
    ]


class SplStepIO(C.Structure):
    """struct spl_step_io (include/splendor_b200.h)."""

    _fields_ = [
        ("actions", C.c_void_p), ("active", C.c_void_p), ("obs", C.c_void_p), ("mask", C.c_void_p),
        ("reward", C.c_void_p), ("terminated", C.c_void_p), ("info", C.c_void_p), ("stats", C.c_void_p),
        ("next_action", C.c_void_p), ("action_key", C.c_uint64), ("action_t", C.c_uint64),
        ("action_t_base", C.c_void_p), ("autoreset", C.c_int32), ("flags", C.c_int32),
        ("obs_f16", C.c_void_p), ("obs_u8", C.c_void_p),
    ]


class SplHostIO(C.Structure):
    """struct spl_host_io (include/splendor_b200.h): HOST pointers (stats: device)."""

    _fields_ = [
        ("actions", C.c_void_p), ("obs", C.c_void_p), ("obs_u8", C.c_void_p), ("mask", C.c_void_p),
        ("reward", C.c_void_p), ("terminated", C.c_void_p), ("info", C.c_void_p), ("next_action", C.c_void_p),
        ("stats", C.c_void_p), ("action_key", C.c_uint64), ("action_t", C.c_uint64),
        ("autoreset", C.c_int32), ("reserved_", C.c_int32),
    ]


class SplendorB200Error(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libsplendor_b200.so and declare the prototypes. Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SplendorB200Error(
            f"CUDA extension not built: {LIB_PATH} is missing. Run `python -m splendor_gym_b200.build` "
            "(needs nvcc; sm_100a). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.spl_init.restype = C.c_int
    L.spl_init.argtypes = []
    L.spl_reset.restype = C.c_int
    L.spl_reset.argtypes = [C.POINTER(SplEnvs), vp, vp, vp, vp, vp]
    L.spl_step.restype = C.c_int
    L.spl_step.argtypes = [C.POINTER(SplEnvs), C.POINTER(SplStepIO), vp]
    L.spl_rollout_random.restype = C.c_int
    L.spl_rollout_random.argtypes = [C.POINTER(SplEnvs), C.POINTER(SplStepIO), C.c_int32, vp]
    L.spl_scripted_action.restype = C.c_int
    L.spl_scripted_action.argtypes = [vp, vp, i64, C.c_int, u64, u64, u64, vp, vp]
    L.spl_masked_sample.restype = C.c_int
    L.spl_masked_sample.argtypes = [vp, vp, i64, C.c_int, u64, u64, u64, vp, vp, vp, vp]
    L.spl_masked_sample_f16.restype = C.c_int
    L.spl_masked_sample_f16.argtypes = [vp, i64, vp, i64, C.c_int, u64, u64, u64, vp, vp, vp, vp]
    L.spl_gae.restype = C.c_int
    L.spl_gae.argtypes = [vp, vp, vp, vp, C.c_int32, i64, C.c_float, C.c_float, vp, vp, vp]
    L.spl_observe.restype = C.c_int
    L.spl_observe.argtypes = [C.POINTER(SplEnvs), vp, vp, vp]
    L.spl_observe_policy.restype = C.c_int
    L.spl_observe_policy.argtypes = [C.POINTER(SplEnvs), vp, vp, vp, vp]
    L.spl_random_action.restype = C.c_int
    L.spl_random_action.argtypes = [vp, i64, u64, u64, u64, vp, vp]
    L.spl_export_state.restype = C.c_int
    L.spl_export_state.argtypes = [C.POINTER(SplEnvs), vp, vp]
    L.spl_import_state.restype = C.c_int
    L.spl_import_state.argtypes = [C.POINTER(SplEnvs), vp, vp, vp]
    L.spl_dual_combine.restype = C.c_int
    L.spl_dual_combine.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, C.c_int, vp]
    L.spl_error_string.restype = C.c_char_p
    L.spl_error_string.argtypes = [C.c_int]
    L.spl_version.restype = C.c_int
    L.spl_host_ret_table.restype = C.c_int
    L.spl_host_ret_table.argtypes = [vp]
    L.spl_launch_count.restype = C.c_int64
    L.spl_timing_enable.restype = C.c_int
    L.spl_timing_enable.argtypes = [C.c_int]
    L.spl_timing_read.restype = C.c_int
    L.spl_timing_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.spl_host_create.restype = C.c_int
    L.spl_host_create.argtypes = [i64, C.c_int32, C.POINTER(vp)]
    L.spl_host_destroy.restype = C.c_int
    L.spl_host_destroy.argtypes = [vp]
    L.spl_host_step.restype = C.c_int
    L.spl_host_step.argtypes = [vp, C.POINTER(SplEnvs), C.POINTER(SplHostIO), vp]
    L.spl_host_observe.restype = C.c_int
    L.spl_host_observe.argtypes = [vp, C.POINTER(SplEnvs), C.POINTER(SplHostIO), vp]
    L.spl_refill_spares.restype = C.c_int
    L.spl_refill_spares.argtypes = [C.POINTER(SplEnvs), vp]
    L.spl_load_deals.restype = C.c_int
    L.spl_load_deals.argtypes = [C.POINTER(SplEnvs), vp, C.c_int32, vp]
    L.spl_rollout_plan.restype = C.c_int
    L.spl_rollout_plan.argtypes = [i64, C.c_int32, C.POINTER(C.c_int32)]
    L.spl_host_expand.restype = C.c_int
    L.spl_host_expand.argtypes = [vp, vp, i64, C.POINTER(SplHostIO)]
    L.spl_host_set_threads.restype = C.c_int
    L.spl_host_set_threads.argtypes = [C.c_int]
    L.spl_host_set_pinning.restype = C.c_int
    L.spl_host_set_pinning.argtypes = [C.c_int]
    L.spl_host_alloc.restype = C.c_int
    L.spl_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.spl_host_free.restype = C.c_int
    L.spl_host_free.argtypes = [vp]
    L.spl_host_get_stats.restype = C.c_int
    L.spl_host_get_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.spl_host_store_rate.restype = C.c_double
    L.spl_host_store_rate.argtypes = [i64, C.c_int, C.c_int]
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().spl_error_string(rc).decode()
        raise SplendorB200Error(f"{what or 'splendor_b200'} failed ({rc}): {msg}")
