"""SplendorVecEnv -- N lock-stepped Splendor games resident in HBM.

Host-side mirror of the reference's vector contract (what ppo_splendor.py:164-173,219-297 consumes from
``gym.vector.SyncVectorEnv`` / the per-env ``dual_step`` loop): observations ``[N,297] int32``, action masks
``[N,45] int8``, rewards ``[N] float32``, terminations ``[N] bool``, same-step auto-reset.  PyTorch only owns
the device buffers and the stream; every computation is a kernel of libsplendor_b200.so reached through
the C ABI (include/splendor_b200.h).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _LazyInfo(dict):
    _DERIVED = {
        "illegal_action": lambda b: (b & L.INFO_ILLEGAL) != 0,
        "draw": lambda b: (b & L.INFO_NOLEGAL_DRAW) != 0,
        "turn_limit": lambda b: (b & L.INFO_TURN_LIMIT) != 0,
        "winner": lambda b: ((b & L.INFO_WINNER_MASK) >> L.INFO_WINNER_SHIFT).to(torch.int8) - 1,
        "error": lambda b: (b & L.INFO_ERROR) != 0,
        "reset": lambda b: (b & L.INFO_RESET) != 0,
    }

    def __init__(self, mask, to_play, bits):
        super().__init__(action_mask=mask, to_play=to_play, info_bits=bits)

    def __missing__(self, key):
        fn = self._DERIVED.get(key)
        if fn is None:
            raise KeyError(key)
        val = fn(self["info_bits"])
        self[key] = val
        return val

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._DERIVED


class SplendorVecEnv:
    """Batched ``SplendorEnv`` (splendor_gym/envs/splendor_env.py:23-130).

    Parameters
    ----------
    num_envs : environments in this shard.
    device : CUDA device.
    seed : base of the per-episode engine-seed schedule; engine seed of (global env g, episode e) is
        ``(seed + 1000003*e + g) mod (2**31-1)`` (the reference draws it from a per-env PCG64 stream,
        envs/splendor_env.py:42-43; explicit engine seeds can be passed to ``reset(seeds=...)``).
    shuffle : ``"mt19937"`` = decks bit-identical to ``initial_state(seed)`` (engine/state.py:181-195);
        ``"philox"`` = native counter-based shuffle (distribution-equivalent, faster).
    env_offset : global index of env 0 (multi-GPU sharding; results do not depend on the GPU count).
    autoreset : same-step auto-reset as in ppo_splendor.py:245-250 (reward/terminated of the finished
        episode, observation/mask of the new one).
    prefetch_deals : with ``shuffle="mt19937"`` and auto-reset, keep the deal of every env's next episode ready
        (``spl_envs_t.spare``) so that a reset does not wait for ``random.Random(seed)`` on the critical path.  ``True`` keeps
        the next 4 episodes ready (384 B per env; they are replaced in batches every 64 lock-steps), an int S the next S (768 B per env for S = 8): ``rollout_random`` then runs whole segments of
        up to 17 (S - 1) + 1 lock-steps in ONE launch with the reference's own decks (a game lasts >= 17 moves).
    obs_format : ``"int32"`` = the reference's observation dtype (envs/splendor_env.py:34-36).  ``"f16"`` = policy-ready:
        ``self.obs_f16`` is an fp16 ``[N, 304]`` tensor (entries 0..296 = the observation, exact; 297..303 = 0) that an MLP
        consumes without the cast of ppo_splendor.py:221 and with a 16-byte-aligned K, and ``self.obs`` holds the same
        values as ``uint8 [N, 297]`` (rollout buffers a quarter of the size).  Auto-reset then needs ``shuffle="philox"`` or
        ``shuffle="mt19937"`` with ``prefetch_deals``.
    """

    num_actions = L.NUM_ACTIONS
    obs_dim = L.OBS_DIM

    def __init__(self, num_envs: int, device="cuda", seed: int = 0, shuffle: str = "philox", env_offset: int = 0,
                 autoreset: bool = True, obs_format: str = "int32", prefetch_deals: bool = True):
        if num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.SplendorB200Error("SplendorVecEnv needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = int(num_envs)
        self.autoreset = bool(autoreset)
        self.shuffle_mode = {"mt19937": L.SHUFFLE_MT19937, "philox": L.SHUFFLE_PHILOX}[shuffle]
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_init(), "spl_init")
        d, n = self.device, self.n
        # structure-of-arrays state in HBM: 4 planes x N x 16 B (hot), N x 96 B deck order (cold)
        self.state = torch.zeros((L.STATE_PLANES, n, 16), dtype=torch.uint8, device=d)
        self.decks = torch.zeros((n, L.DECK_STRIDE), dtype=torch.uint8, device=d)
        self.episode = torch.zeros(n, dtype=torch.int32, device=d)
        self.scratch = torch.zeros(n + 4, dtype=torch.int32, device=d)
        if obs_format not in ("int32", "f16"):
            raise ValueError("obs_format must be 'int32' or 'f16'")
        self.obs_format = obs_format
        if obs_format == "f16":
            if self.autoreset and self.shuffle_mode != L.SHUFFLE_PHILOX and not prefetch_deals:
                raise L.SplendorB200Error("obs_format='f16' with auto-reset needs shuffle='philox' or prefetched deals")
            self.obs = torch.zeros((n, L.OBS_DIM), dtype=torch.uint8, device=d)
            self.obs_f16 = torch.zeros((n, L.OBS_F16_PITCH), dtype=torch.float16, device=d)
        else:
            self.obs = torch.zeros((n, L.OBS_DIM), dtype=torch.int32, device=d)
            self.obs_f16 = None
        self.mask = torch.zeros((n, L.NUM_ACTIONS), dtype=torch.int8, device=d)
        self.reward = torch.zeros(n, dtype=torch.float32, device=d)
        self._terminated = torch.zeros(n, dtype=torch.uint8, device=d)
        self.terminated = self._terminated.view(torch.bool)
        self.truncated = torch.zeros(n, dtype=torch.bool, device=d)
        self.info_bits = torch.zeros(n, dtype=torch.uint8, device=d)
        self.stats = torch.zeros(8, dtype=torch.int64, device=d)
        self.next_action = torch.zeros(n, dtype=torch.int32, device=d)
        # bit-exact decks with auto-reset: prefetched deal of every env's next episode + refill list (struct spl_envs.spare)
        self.spare = None
        self.spare_slots = 0
        if self.shuffle_mode == L.SHUFFLE_MT19937 and self.autoreset and prefetch_deals:
            self.spare_slots = 4 if prefetch_deals is True else max(1, min(int(prefetch_deals), L.MAX_SPARE_SLOTS))
            self.spare = torch.zeros(n * self.spare_slots * L.DECK_STRIDE + (n * self.spare_slots + 4) * 4, dtype=torch.uint8, device=d)
        self._envs = L.SplEnvs(
            state=self.state.data_ptr(), decks=self.decks.data_ptr(), episode=self.episode.data_ptr(),
            scratch=self.scratch.data_ptr(), stride=n, n=n, env_offset=int(env_offset), seed_base=int(seed),
            shuffle_mode=self.shuffle_mode, spare_slots=self.spare_slots, spare=None if self.spare is None else self.spare.data_ptr(),
        )
        self._envs_ref = C.byref(self._envs)
        self._device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        # gymnasium.vector-style attributes (what gym.vector.SyncVectorEnv exposes, ppo_splendor.py:151-159)
        from .envs._gym_compat import spaces
        import numpy as _np

        self.num_envs = n
        self.single_action_space = spaces.Discrete(L.NUM_ACTIONS)
        self.single_observation_space = spaces.Box(low=0, high=50, shape=(L.OBS_DIM,), dtype=_np.int32)
        self._io = L.SplStepIO()
        self._t = 0  # lock-step counter (drives the Philox action stream)
        self.t_base = None  # optional device int64 scalar added to the counter (CUDA-graph replays)
        self.action_key = 0xB200
        self._is_reset = False

    def close(self) -> None:
        """Releases the host-path staging context, if any (the tensors go with the object)."""
        self.close_host()

    def episode_statistics(self) -> Dict[str, int]:
        """Counters accumulated by the step kernels since construction (one host read)."""
        return dict(zip(L.STAT_NAMES, self.stats.cpu().tolist()))

    # ------------------------------------------------------------------ plumbing
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def info(self) -> Dict[str, torch.Tensor]:
        """The reference's info dict, batched (envs/splendor_env.py:47-48,82-88).  "action_mask", "to_play" and
        "info_bits" are views of the step outputs; the boolean / winner entries are decoded from the info bits on
        first access (no extra kernels on the step path unless somebody asks)."""
        return _LazyInfo(self.mask, self.obs[:, 294], self.info_bits)

    @staticmethod
    def final_rewards(info_bits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``SplendorEnv.get_final_rewards`` (envs/splendor_env.py:92-115) from info bits -> (r0, r1, present)."""
        b = info_bits.to(torch.int32)
        present = ((b & L.INFO_TERMINATED) != 0) & ((b & (L.INFO_NOLEGAL_DRAW | L.INFO_ERROR)) == 0)
        w = ((b & L.INFO_WINNER_MASK) >> L.INFO_WINNER_SHIFT) - 1
        draw = torch.where((b & L.INFO_TURN_LIMIT) != 0, -0.1, 0.0).to(torch.float32)
        r0 = torch.where(w < 0, draw, torch.where(w == 0, 1.0, -1.0).to(torch.float32))
        r1 = torch.where(w < 0, draw, torch.where(w == 1, 1.0, -1.0).to(torch.float32))
        zero = torch.zeros_like(r0)
        return torch.where(present, r0, zero), torch.where(present, r1, zero), present

    # ------------------------------------------------------------------ reset / step
    def reset(self, *, seed: Optional[int] = None, seeds: Optional[torch.Tensor] = None,
              reset_mask: Optional[torch.Tensor] = None, options=None):
        """``SplendorEnv.reset`` for all envs (or those flagged in ``reset_mask``) -> (obs, info)."""
        if seed is not None:
            if reset_mask is not None and int(seed) != int(self._envs.seed_base):
                # the prefetched deals of the envs that are NOT reset were computed from the current base seed
                raise ValueError("reset(seed=...) re-seeds every env: it cannot be combined with reset_mask")
            self._envs.seed_base = int(seed)
        if seeds is not None:
            seeds = seeds.to(device=self.device, dtype=torch.int64).contiguous()
        if reset_mask is not None:
            reset_mask = reset_mask.to(device=self.device).view(-1).to(torch.uint8).contiguous()
        f16 = self.obs_format == "f16"
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_reset(C.byref(self._envs), _ptr(seeds), _ptr(reset_mask), None if f16 else self.obs.data_ptr(),
                                       None if f16 else self.mask.data_ptr(), self._stream()), "spl_reset")
            if f16:
                L.check(self.lib.spl_observe_policy(C.byref(self._envs), self.obs_f16.data_ptr(), self.obs.data_ptr(),
                                                    self.mask.data_ptr(), self._stream()), "spl_observe_policy")
        if reset_mask is None:
            self._t = 0
        self._is_reset = True
        return self.obs, {"action_mask": self.mask, "to_play": self.obs[:, 294]}

    def step(self, actions: torch.Tensor, *, active: Optional[torch.Tensor] = None, out_obs: Optional[torch.Tensor] = None,
             out_mask: Optional[torch.Tensor] = None, sample_next: bool = False, autoreset: Optional[bool] = None,
             out_reward: Optional[torch.Tensor] = None, out_terminated: Optional[torch.Tensor] = None,
             out_next_action: Optional[torch.Tensor] = None, write_obs: bool = True, out_obs_f16: Optional[torch.Tensor] = None,
             check: bool = False):
        """``SplendorEnv.step`` for every env in lock-step -> (obs, reward, terminated, truncated, info).

        Where the reference raises, the batched step records ``SPL_INFO_ERROR`` in ``info_bits`` and leaves that env
        untouched; ``check=True`` (one device->host read of the info bytes) raises like the reference does for the
        first offending env: ``ValueError`` for an action outside ``[0, 45)`` (envs/splendor_env.py:62-63),
        ``RuntimeError`` for a step after termination (:53-54, only possible without auto-reset).

        ``out_obs`` / ``out_mask`` redirect the observation / mask of this step into caller storage (e.g. a
        rollout buffer slice).  ``sample_next`` also draws a uniform random legal action for the returned
        mask into ``self.next_action`` (fused; same stream as ``sample_random_actions``).
        """
        assert self._is_reset, "Call reset() first"
        if actions.dtype != torch.int32 or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if active is not None:
            active = active.to(device=self.device).view(-1).to(torch.uint8).contiguous()
        obs = self.obs if out_obs is None else out_obs
        mask = self.mask if out_mask is None else out_mask
        io = self._io
        io.actions, io.active = actions.data_ptr(), _ptr(active)
        reward = self.reward if out_reward is None else out_reward
        term = self._terminated if out_terminated is None else out_terminated
        nxt = self.next_action if out_next_action is None else out_next_action
        if self.obs_format == "f16":  # fp16 policy input + byte observation instead of the int32 array
            f16 = self.obs_f16 if out_obs_f16 is None else out_obs_f16
            io.obs, io.mask = None, mask.data_ptr()
            io.obs_f16, io.obs_u8 = (f16.data_ptr() if write_obs else None), (obs.data_ptr() if write_obs else None)
        else:
            io.obs, io.mask = (obs.data_ptr() if write_obs else None), mask.data_ptr()
            io.obs_f16, io.obs_u8 = None, None
        io.reward, io.terminated, io.info = reward.data_ptr(), term.data_ptr(), self.info_bits.data_ptr()
        io.stats = self.stats.data_ptr()
        io.next_action = nxt.data_ptr() if (sample_next or out_next_action is not None) else None
        io.action_key, io.action_t = self.action_key, self._t + 1
        io.action_t_base = _ptr(self.t_base)
        io.autoreset = int(self.autoreset if autoreset is None else autoreset)
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_step(C.byref(self._envs), C.byref(io), self._stream()), "spl_step")
        self._t += 1
        if check:
            bad = torch.nonzero(self.info_bits & L.INFO_ERROR).flatten()
            if bad.numel():
                i = int(bad[0])
                if int(self.info_bits[i]) & L.INFO_TERMINATED:
                    raise RuntimeError(f"env {i}: step() called on a terminal state ({bad.numel()} envs in error); call reset()")
                raise ValueError(f"env {i}: action {int(actions[i])} out of range ({bad.numel()} envs in error)")
        if out_reward is not None or out_terminated is not None:
            return obs, reward, term.view(torch.bool), self.truncated, None
        return obs, self.reward, self.terminated, self.truncated, self.info()

    def rollout_random(self, steps: int, actions: torch.Tensor, *, obs: Optional[torch.Tensor], mask: Optional[torch.Tensor],
                       reward: torch.Tensor, terminated: torch.Tensor, next_actions: torch.Tensor,
                       info: Optional[torch.Tensor] = None, refill: bool = True) -> None:
        """``steps`` lock-steps of uniform-random-legal play (scripts/random_rollout.py:13-30, batched) with same-step
        auto-reset in ONE kernel launch, streamed into step-major rollout buffers: ``obs [steps,N,297]``,
        ``mask [steps,N,45]``, ``reward / terminated / info [steps,N]``, ``next_actions [steps+1,N]`` (row t+1 = the
        action sampled after step t; row 0 is not written).  ``actions [N]`` are the actions of the first step.
        Bit-identical to ``steps`` calls of ``step(..., sample_next=True)``.  Needs autoreset and ``shuffle="philox"`` or
        ``shuffle="mt19937"`` with prefetched deals (``prefetch_deals=8`` for segments of up to 120 lock-steps).
        ``refill=False`` (MT19937): the launch does not replace the deals it takes; the caller runs ``refill_deals()``,
        typically on a side stream while the NEXT segment is already running (``prefetch_deals=16`` then covers the two
        segments in flight), and makes launch k+2 wait for the refill that followed launch k."""
        assert self._is_reset, "Call reset() first"
        n = self.n
        assert actions.dtype == torch.int32 and actions.is_contiguous() and actions.numel() == n
        assert next_actions.dtype == torch.int32 and next_actions.is_contiguous() and next_actions.numel() == (steps + 1) * n
        assert reward.numel() == steps * n and terminated.numel() == steps * n
        io = self._io
        io.actions, io.active = actions.data_ptr(), None
        io.obs_f16, io.obs_u8 = None, None
        io.obs, io.mask = _ptr(obs), _ptr(mask)
        io.reward, io.terminated, io.info = reward.data_ptr(), terminated.data_ptr(), _ptr(info)
        io.stats = self.stats.data_ptr()
        io.next_action = next_actions.data_ptr()
        io.action_key, io.action_t = self.action_key, self._t + 1
        io.action_t_base = _ptr(self.t_base)
        io.autoreset = 1
        io.flags = 0 if refill else L.IO_ASYNC_REFILL
        try:
            with torch.cuda.device(self.device):
                L.check(self.lib.spl_rollout_random(C.byref(self._envs), C.byref(io), int(steps), self._stream()), "spl_rollout_random")
        finally:
            io.flags = 0
        self._t += steps

    def refill_deals(self) -> None:
        """Replace every prefetched deal taken since the last refill (``spl_refill_spares``).  ``step`` / ``step_host`` /
        ``rollout_random`` do this on their own cadence; a caller that replays a captured single-step CUDA graph (whose
        lock-step counter is frozen) calls it every <= 16 x ``spare_slots`` lock-steps.  Runs on the current torch stream:
        inside ``with torch.cuda.stream(side)`` it overlaps ``rollout_random(..., refill=False)`` launches on the main one."""
        if self.spare is None:
            raise L.SplendorB200Error("refill_deals needs shuffle='mt19937' with prefetch_deals")
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_refill_spares(C.byref(self._envs), self._stream()), "spl_refill_spares")

    # ------------------------------------------------------------------ host-buffer API (NumPy-side callers)
    def _host_setup(self, obs_dtype):
        """Host result arrays (reused every call, like gym.vector's pre-allocated observation buffers) + the library's
        staging context.  ``obs_dtype``: ``torch.int32`` = the reference's dtype (envs/splendor_env.py:34-36);
        ``torch.uint8`` = the same values as bytes (a quarter of the host memory traffic; all entries are < 256)."""
        H = getattr(self, "_host", None)
        if H is not None and H["obs"].dtype == obs_dtype:
            return H
        n = self.n
        if H is None:
            handle = C.c_void_p()
            with torch.cuda.device(self.device):
                L.check(self.lib.spl_host_create(n, 0, C.byref(handle)), "spl_host_create")
            H = dict(handle=handle, blocks=[], obs_by_dtype={}, plans={})
            self._host = H
            for name, shape, dt in (("mask", (n, L.NUM_ACTIONS), torch.int8), ("reward", (n,), torch.float32),
                                    ("terminated", (n,), torch.uint8), ("info_bits", (n,), torch.uint8),
                                    ("next_action", (n,), torch.int32)):
                H[name] = self._host_array(H, shape, dt)
            H["truncated"] = torch.zeros(n, dtype=torch.bool)
        if obs_dtype not in H["obs_by_dtype"]:
            H["obs_by_dtype"][obs_dtype] = self._host_array(H, (n, L.OBS_DIM), obs_dtype)
        H["obs"] = H["obs_by_dtype"][obs_dtype]
        return H

    def _host_array(self, H, shape, dtype) -> torch.Tensor:
        """A zeroed CPU tensor over ``spl_host_alloc`` memory (pinned + mapped, huge pages requested): the GPU can write
        a share of the results into it directly.  Plain host memory to the caller (``.numpy()`` works)."""
        import numpy as np

        count = 1
        for d in shape:
            count *= int(d)
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_host_alloc(max(nbytes, 1), C.byref(ptr)), "spl_host_alloc")
        H["blocks"].append(ptr)
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8)).view(dtype).view(*shape)

    def _host_plan(self, obs_dtype, sample_next):
        """Everything about a host call that does not change from step to step, built once per (dtype, sample_next): the
        filled ``spl_host_io`` and the result tuple (views of the persistent arrays).  The per-step Python cost of
        ``step_host`` is then three field writes and one foreign call."""
        H = self._host_setup(obs_dtype)
        plan = H["plans"].get((obs_dtype, sample_next))
        if plan is None:
            io = L.SplHostIO()
            obs = H["obs"]
            io.obs = obs.data_ptr() if obs_dtype == torch.int32 else None
            io.obs_u8 = obs.data_ptr() if obs_dtype == torch.uint8 else None
            io.mask, io.reward = H["mask"].data_ptr(), H["reward"].data_ptr()
            io.terminated, io.info = H["terminated"].data_ptr(), H["info_bits"].data_ptr()
            io.next_action = H["next_action"].data_ptr() if sample_next else None
            io.stats = self.stats.data_ptr()
            io.action_key = self.action_key
            info = {"action_mask": H["mask"], "to_play": obs[:, 294], "info_bits": H["info_bits"]}
            if sample_next:
                info["next_action"] = H["next_action"]
            plan = dict(io=io, io_ref=C.byref(io), obs=obs, info=info,
                        step_out=(obs, H["reward"], H["terminated"].view(torch.bool), H["truncated"], info))
            H["plans"][(obs_dtype, sample_next)] = plan
        H["obs"] = plan["obs"]
        return H, plan

    def _host_call(self, fn, name, plan, H, actions_ptr, action_t, autoreset):
        io = plan["io"]
        io.actions = actions_ptr
        io.action_t = action_t
        io.autoreset = int(autoreset)
        if torch.cuda.current_device() == self._device_index:
            rc = fn(H["handle"], self._envs_ref, plan["io_ref"], torch._C._cuda_getCurrentRawStream(self._device_index))
        else:
            with torch.cuda.device(self.device):
                rc = fn(H["handle"], self._envs_ref, plan["io_ref"], self._stream())
        if rc != 0:
            L.check(rc, name)

    def reset_host(self, *, obs_dtype=torch.int32, sample_next: bool = False, **kw):
        """``reset`` returning HOST tensors -> (obs [N,297], {"action_mask": int8 [N,45], "to_play": ...})."""
        if obs_dtype not in (torch.int32, torch.uint8):
            raise ValueError("obs_dtype must be torch.int32 or torch.uint8")
        self.reset(**kw)
        H, plan = self._host_plan(obs_dtype, sample_next)
        self._host_call(self.lib.spl_host_observe, "spl_host_observe", plan, H, None, self._t, False)
        return plan["obs"], {"action_mask": H["mask"], "to_play": plan["obs"][:, 294]}

    def step_host(self, actions, *, obs_dtype=torch.int32, sample_next: bool = False, autoreset: Optional[bool] = None):
        """``SplendorEnv.step`` for every env with HOST arrays in and out -- the call a NumPy-side vector loop makes
        (ppo_splendor.py:235-285).  ``actions``: int32 ``numpy.ndarray`` / CPU tensor ``[N]``.  Returns CPU tensors
        ``(obs, reward, terminated, truncated, info)`` owned by the env and overwritten by the next call;
        ``info["next_action"]`` (with ``sample_next``) is a uniform random legal action per env for the new mask.
        Device work: H2D actions -> step kernel (compact outputs) -> push kernel, which stores 64-env groups over PCIe
        into pinned host memory: nibble-packed for the share that pinned host threads widen into the reference-typed
        arrays, already widened for the rest (csrc/spl_host.cu).  Values are those of ``step``."""
        assert self._is_reset, "Call reset() first"
        if obs_dtype is not torch.int32 and obs_dtype is not torch.uint8:
            raise ValueError("obs_dtype must be torch.int32 or torch.uint8")
        autoreset = self.autoreset if autoreset is None else autoreset
        if autoreset and self.shuffle_mode != L.SHUFFLE_PHILOX and self.spare is None:
            raise L.SplendorB200Error("step_host with auto-reset needs shuffle='philox' or shuffle='mt19937' with prefetch_deals "
                                      "(resets must happen inside the step kernel)")
        if isinstance(actions, torch.Tensor):
            a = actions
            if a.dtype != torch.int32 or not a.is_contiguous() or a.device.type != "cpu":
                a = a.to(device="cpu", dtype=torch.int32).contiguous()
            if a.numel() != self.n:
                raise ValueError("actions must have one entry per env")
            ptr = a.data_ptr()
        else:
            import numpy as np

            a = actions
            if not (isinstance(a, np.ndarray) and a.dtype == np.int32 and a.flags.c_contiguous):
                a = np.ascontiguousarray(actions, dtype=np.int32)
            if a.size != self.n:
                raise ValueError("actions must have one entry per env")
            ptr = a.__array_interface__["data"][0]
        H, plan = self._host_plan(obs_dtype, sample_next)
        self._host_call(self.lib.spl_host_step, "spl_host_step", plan, H, ptr, self._t + 1, autoreset)
        self._t += 1
        return plan["step_out"]

    def close_host(self) -> None:
        H = getattr(self, "_host", None)
        if H is not None:
            self.lib.spl_host_destroy(H["handle"])
            for k in [k for k, v in H.items() if isinstance(v, torch.Tensor)]:
                del H[k]  # views over the blocks freed below
            for ptr in H["blocks"]:
                self.lib.spl_host_free(ptr)
            self._host = None

    def host_stats(self) -> dict:
        """Timings of the last ``step_host`` / ``reset_host`` (``spl_host_get_stats``), microseconds from the start of the call."""
        out = (C.c_double * 8)()
        L.check(self.lib.spl_host_get_stats(self._host["handle"], out), "spl_host_get_stats")
        keys = ("call_us", "enqueued_us", "first_group_us", "workers_done_us", "gpu_share_done_us", "gpu_written_share", "threads", "gpu_writable")
        return dict(zip(keys, [float(v) for v in out]))

    def observe(self, out_obs: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None):
        """encode_observation + legal_moves of the current states (no step)."""
        obs = self.obs if out_obs is None else out_obs
        mask = self.mask if out_mask is None else out_mask
        with torch.cuda.device(self.device):
            if self.obs_format == "f16":
                L.check(self.lib.spl_observe_policy(C.byref(self._envs), self.obs_f16.data_ptr(), obs.data_ptr(), mask.data_ptr(),
                                                    self._stream()), "spl_observe_policy")
            else:
                L.check(self.lib.spl_observe(C.byref(self._envs), obs.data_ptr(), mask.data_ptr(), self._stream()), "spl_observe")
        return obs, mask

    def sample_random_actions(self, mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``random_opponent`` (wrappers/selfplay.py:66-73) for every env: uniform over the legal actions."""
        mask = self.mask if mask is None else mask
        out = self.next_action if out is None else out
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_random_action(mask.data_ptr(), self.n, self._envs.env_offset, self.action_key, self._t,
                                               out.data_ptr(), self._stream()), "spl_random_action")
        return out

    # ------------------------------------------------------------------ state exchange (tests, debugging)
    def export_state(self) -> torch.Tensor:
        rows = torch.empty((self.n, L.ROW_LEN), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_export_state(C.byref(self._envs), rows.data_ptr(), self._stream()), "spl_export_state")
        return rows

    def import_state(self, rows: torch.Tensor, which: Optional[torch.Tensor] = None) -> None:
        rows = rows.to(device=self.device, dtype=torch.int32).contiguous()
        assert rows.shape == (self.n, L.ROW_LEN)
        if which is not None:
            which = which.to(device=self.device).view(-1).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_import_state(C.byref(self._envs), rows.data_ptr(), _ptr(which), self._stream()), "spl_import_state")
        self._is_reset = True

    # ------------------------------------------------------------------ replay of a reference run's decks
    def set_episode_seeds(self, table: Optional[torch.Tensor]) -> None:
        """Engine seeds of every env's episodes 1..E (``[N, E]`` integers; ``None`` = back to the library's schedule), for
        ``shuffle="mt19937"``.  The reference's ``SplendorEnv`` draws a fresh engine seed from its own PCG64 stream on every
        reset (envs/splendor_env.py:42-43; the vector env re-draws on each auto-reset, ppo_splendor.py:246-247): pass those
        draws here and ``reset(seeds=first_draws)``, and every auto-reset deals exactly the reference's decks.
        Call it BEFORE ``reset`` (which fills the ring of prefetched deals from it)."""
        if table is None:
            self._episode_seeds = None
            self._envs.episode_seeds, self._envs.episode_seed_count = None, 0
            return
        if self.shuffle_mode != L.SHUFFLE_MT19937:
            raise L.SplendorB200Error("set_episode_seeds needs shuffle='mt19937' (the reference's random.Random decks)")
        t = torch.as_tensor(table).to(device=self.device, dtype=torch.int64).contiguous()
        if t.dim() != 2 or t.shape[0] != self.n or t.shape[1] < 1:
            raise ValueError("episode seeds must be [N, E]")
        self._episode_seeds = t  # (kept alive: the struct holds its address)
        self._envs.episode_seeds, self._envs.episode_seed_count = t.data_ptr(), int(t.shape[1])

    def load_deals(self, deals) -> None:
        """Caller-supplied deck permutations for every env's next ``k`` episodes (``spl_load_deals``): ``deals`` is
        ``uint8 [N, k, 96]`` (host or device), ``k <= prefetch_deals``; row format in include/splendor_b200.h, or
        ``deals_from_rows(rows)`` from exported initial-state rows (e.g. of a reference run)."""
        if self.spare is None:
            raise L.SplendorB200Error("load_deals needs shuffle='mt19937' with prefetch_deals")
        d = torch.as_tensor(deals)
        if d.dtype != torch.uint8 or d.dim() != 3 or d.shape[0] != self.n or d.shape[2] != L.DECK_STRIDE:
            raise ValueError("deals must be uint8 [N, k, 96]")
        d = d.contiguous()
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_load_deals(C.byref(self._envs), d.data_ptr(), int(d.shape[1]), self._stream()), "spl_load_deals")

    @staticmethod
    def deals_from_rows(rows: torch.Tensor) -> torch.Tensor:
        """Flat rows of INITIAL states (``export_state`` layout, as ``initial_state`` leaves them: 36 / 26 / 16 cards in the
        decks, 12 on the board) -> ``uint8 [N, 96]`` deal rows: the shuffled tier lists with the four board cards put back
        at the end in dealing order (engine/state.py:190-191) and the three visible nobles."""
        r = torch.as_tensor(rows).to(torch.int64)
        out = torch.full((r.shape[0], L.DECK_STRIDE), 255, dtype=torch.int64, device=r.device)
        for first, size, deck_col, board_col in ((0, 40, 76, 52), (40, 30, 116, 56), (70, 20, 146, 60)):
            out[:, first:first + size - 4] = r[:, deck_col:deck_col + size - 4]
            out[:, first + size - 4:first + size] = r[:, board_col:board_col + 4].flip(1)
        out[:, 90:93] = r[:, 67:70]
        return out.to(torch.uint8)

    # ------------------------------------------------------------------ dual step (self-play turn)
    def dual_step(self, agent_actions: torch.Tensor, opponent_policy: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                  reward_mode: str = "native"):
        """``DualStepNativeWrapper.dual_step`` (wrappers/dual_step_native.py:90-193), batched.

        Phase 1: every env applies the agent's (player 0) action.  Phase 2: envs still running apply
        ``opponent_policy(obs, mask) -> actions``.  Returns
        ``(agent_obs, agent_reward, opp_obs, opp_reward, done, info)``; ``agent_obs is opp_obs`` like the
        reference (:182-191).  An env whose agent action was rejected (illegal) skips phase 2 (the
        reference raises there, :149-150).
        """
        if not hasattr(self, "_dual"):
            d, n = self.device, self.n
            self._dual = dict(
                r1=torch.zeros(n, dtype=torch.float32, device=d), t1=torch.zeros(n, dtype=torch.uint8, device=d),
                i1=torch.zeros(n, dtype=torch.uint8, device=d), agent_r=torch.zeros(n, dtype=torch.float32, device=d),
                opp_r=torch.zeros(n, dtype=torch.float32, device=d), done=torch.zeros(n, dtype=torch.uint8, device=d),
            )
        D = self._dual
        obs, _, _, _, _ = self.step(agent_actions)
        D["r1"].copy_(self.reward)
        D["t1"].copy_(self._terminated)
        D["i1"].copy_(self.info_bits)
        active = (D["t1"] == 0) & ((D["i1"] & (L.INFO_ILLEGAL | L.INFO_ERROR)) == 0)
        opp_actions = opponent_policy(self.obs, self.mask)
        obs, _, _, _, _ = self.step(opp_actions, active=active)
        with torch.cuda.device(self.device):
            L.check(self.lib.spl_dual_combine(D["r1"].data_ptr(), D["t1"].data_ptr(), D["i1"].data_ptr(), self.reward.data_ptr(),
                                              self._terminated.data_ptr(), self.info_bits.data_ptr(), self.n,
                                              D["agent_r"].data_ptr(), D["opp_r"].data_ptr(), D["done"].data_ptr(),
                                              {"native": 0, "selfplay": 1}[reward_mode], self._stream()), "spl_dual_combine")
        info = {"action_mask": self.mask, "to_play": self.obs[:, 294], "info_bits_agent": D["i1"], "info_bits_opponent": self.info_bits}
        return obs, D["agent_r"], obs, D["opp_r"], D["done"].view(torch.bool), info
