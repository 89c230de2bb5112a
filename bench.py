#!/usr/bin/env python
"""bench.py -- env-steps/s of the hot path (legality check + step + legal mask + observation encode +
same-step auto-reset) on N B200s, with the roofline of the dominant kernel, a CPU baseline, and the
end-to-end number through the public API with host buffers.

Workload (BASELINE.json configs[1]): 2-player random-legal-policy lock-step rollout, 65,536 envs per GPU.
One bench "step" = one rollout segment: ROLLOUT lock-steps of all envs, each writing its observations /
masks / rewards / terminations / actions into a [ROLLOUT, N, ...] rollout buffer (what ppo_splendor.py's
obs_buf/masks_buf/... hold, ppo_splendor.py:210-297), replayed as one CUDA graph.  The rollout buffer
(10.4 GB at the defaults) is far larger than the 126 MB L2, so no L2 flush is needed between iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--rollout T] [--impl reference]
  torchrun ... bench.py --gpus N ...      (one rank per GPU; env shards are independent, weak scaling)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_FULL = 1370  # algorithmic bytes per env-step: obs 1188 + mask 45 + reward 4 + term 1 + action 4 + state 64 r + 64 w
B_MASKSTEP = 182
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def workload_config(N: int, T: int, write_obs: bool = True) -> dict:
    """`config` of the JSON line: the workload only, so that both arms (--impl b200 / reference) print the same object."""
    per_step_bytes = N * ((1188 if write_obs else 0) + 45 + 4 + 1 + 4)
    return {
        "workload": ("2-player random-legal lock-step rollout, %d envs per GPU (BASELINE configs[1]%s), step+mask+obs+same-step auto-reset"
                     % (N, "" if N == 65536 else "; envs overridden")) if write_obs else
                    "simplified take-3 rules, mask+step only (BASELINE configs[3]), %d envs per GPU" % N,
        "envs_per_gpu": N, "lock_steps_per_step": T,
        "l2": "results of one step (%.1f GB per GPU: [lock_steps, envs, ...] rollout buffers) are larger than the 126 MB L2; no flush"
              % (T * per_step_bytes / 1e9),
    }


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--rollout", type=int, default=128, help="lock-steps per bench step (ppo_splendor.py --num-steps)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-obs", action="store_true", help="config 4: mask + step only (no observation encode)")
    ap.add_argument("--shuffle", default="philox", choices=["philox", "mt19937"])
    ap.add_argument("--deal-slots", type=int, default=8, help="shuffle=mt19937, rollout mode: prefetched deals kept per env")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "lockstep"],
                    help="rollout: one persistent kernel per segment; lockstep: one spl_step launch per lock-step")
    ap.add_argument("--skip-lockstep", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="do not run the bounded sub-results for BASELINE configs 3 / 4 / 5")
    ap.add_argument("--stats-every", type=int, default=1, help="multi-GPU: all-reduce the episode statistics every M bench steps")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML from a thread, ~1 kHz; nvidia-smi fallback)."""

    def __init__(self, index: int):
        import threading

        self.samples, self.maxes, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._thread = None
        self._smi = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            uuid = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(index)
            self._nv, self._h = nv, h
            self._max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            try:
                self._f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                             stdout=self._f, stderr=subprocess.DEVNULL)
            except Exception:
                self._smi = None

    def _loop(self):
        nv, h = self._nv, self._h
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.001)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if self.samples:
                return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self._max, "reasons": sorted(self.reasons),
                        "samples": len(self.samples), "source": "nvml"}
        if self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            self._f.flush()
            self._f.seek(0)
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self._f.read().splitlines():
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            if sm:
                return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ------------------------------------------------------------------------------------------------- CPU legs
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_rollout(envs: int, seconds: float, threads: int | None = None):
    """Oracle port of the reference engine on the host cores: random-legal lock-step rollout with same-step
    auto-reset, full step+mask+obs per env-step. Returns (steps_per_s, threads, description)."""
    from oracle import oracle as O

    O.build()
    O.set_num_threads(threads or host_threads())  # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    nthreads = O.num_threads()
    v = O.OracleVec(envs, seed_base=0)
    v.reset()
    v.rollout_random(0xB200, 0, 2)  # warm-up
    t0 = time.perf_counter()
    done, T, chunk = 0, 2, 4
    while True:
        done += v.rollout_random(0xB200, T, chunk)
        T += chunk
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, nthreads, f"{envs} envs x {T - 2} lock-steps ({done} env-steps, {el:.1f} s), C port of the reference engine (oracle/), OpenMP"


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; the Python reference
    cannot travel to the GPU box), all threads, same workload/metric. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as O

    O.build()
    O.set_num_threads(host_threads())
    nthreads = O.num_threads()
    envs = args.envs
    per_step_bytes = envs * (1188 + 45 + 4 + 1 + 4)
    T_seg = max(1, min(args.rollout, int(40e9 // per_step_bytes)))  # (the GPU arm's cap on the rollout buffer)
    v = O.OracleVec(envs, seed_base=0)
    v.reset()
    # one bench step = the GPU arm's step: one segment of T_seg lock-steps of all envs (~0.8 s on 16 cores at the defaults).
    # Warm-up steps are shortened to 8 lock-steps (they only have to touch the pages and spin up the threads).
    T = 0
    for _ in range(max(args.warmup, 1)):
        v.rollout_random(0xB200, T, 8)
        T += 8
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += v.rollout_random(0xB200, T, T_seg)
        T += T_seg
    el = time.perf_counter() - t0
    val = done / el
    sample = f"{envs} envs x {T_seg} lock-steps per step x {args.steps} steps ({done} env-steps, {el:.1f} s), oracle C port of the reference engine, {nthreads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": workload_config(envs, T_seg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample,
                         "python_reference": python_reference_record()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def python_reference_record():
    """Throughput of the UNMODIFIED Python engine (north_star's named baseline: SplendorEnv under gymnasium's Sync / Async
    vector env), measured by oracle/time_pyref.py in the build container -- /root/reference cannot travel to the GPU box --
    and committed as tests/golden/pyref_throughput.json.  Another box: shown next to, never instead of, the same-box port."""
    p = os.path.join(ROOT, "tests", "golden", "pyref_throughput.json")
    try:
        with open(p) as f:
            rec = json.load(f)
        rec["note"] = "measured in the build container (other box), see 'host'; the port above ran on this box"
        return rec
    except Exception:
        return None



def extra_configs(args, dev, rank, world, lib, peak, peak_src):
    """Bounded sub-results for BASELINE configs[2] (PPO-MLP in the loop, 262,144 envs), configs[3] (mask+step only, 1M envs)
    and configs[4] (1,048,576 envs per GPU x 30 lock-steps, random policy; all ranks, stats all-reduced over NCCL)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200 import _lib as L

    out = {}

    def rollout_case(n, T, write_obs, reps):
        env = SplendorVecEnv(n, device=dev, seed=20261018, shuffle="philox", env_offset=rank * n, autoreset=True)
        obs = torch.zeros((T, n, 297), dtype=torch.int32, device=dev) if write_obs else None
        mask = torch.zeros((T, n, 45), dtype=torch.int8, device=dev)
        rew = torch.zeros((T, n), dtype=torch.float32, device=dev)
        term = torch.zeros((T, n), dtype=torch.uint8, device=dev)
        act = torch.zeros((T + 1, n), dtype=torch.int32, device=dev)
        env.t_base = torch.zeros(1, dtype=torch.int64, device=dev)
        env.reset()
        env.sample_random_actions(out=act[0])
        stats = torch.zeros(8, dtype=torch.int64, device=dev)

        def seg():
            env._t = 0
            env.rollout_random(T, act[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=act)
            act[0].copy_(act[T])
            env.t_base += T

        for _ in range(3):
            seg()
        torch.cuda.synchronize()
        L.check(lib.spl_timing_enable(1))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            seg()
        if world > 1:  # the episode statistics are the only cross-GPU traffic
            stats.copy_(env.stats)
            dist.all_reduce(stats)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tot, cnt = C.c_double(), C.c_int64()
        L.check(lib.spl_timing_read(C.byref(tot), C.byref(cnt)))
        L.check(lib.spl_timing_enable(0))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        plan = (C.c_int32 * 6)()
        L.check(lib.spl_rollout_plan(n, T, plan))
        ob = (1188 if write_obs else 0) + 45 + 4 + 1 + 4
        bytes_per_launch = n * T * ob + n * 132 * int(plan[3])
        k_ms = tot.value / max(1, cnt.value)
        ach = bytes_per_launch / (k_ms * 1e-3) / 1e9
        sv = B_FULL if write_obs else B_MASKSTEP
        res = {"value": n * T * reps * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "envs_per_gpu": n, "lock_steps_per_launch": T,
               "launches": reps, "ms_per_launch": ms / reps,
               "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "kernel": "spl_rollout_kernel",
                            "kernel_ms": k_ms, "algorithmic_bytes_per_env_step": bytes_per_launch / (n * T),
                            "survey_8d_bytes_per_env_step": sv, "frac_with_survey_8d_bytes": n * T * sv / (k_ms * 1e-3) / 1e9 / peak,
                            "peak_source": peak_src},
               "episodes": int(stats[0].item()) if world > 1 else int(env.stats[0].item())}
        env.close()
        return res

    try:
        r = rollout_case(1 << 20, 30, False, 10)
        r["workload"] = "BASELINE configs[3]: mask+step only (no observation encode) at 1,048,576 envs per GPU, random-legal policy, same-step auto-reset"
        out["config4"] = r
        torch.cuda.empty_cache()
        r = rollout_case(1 << 20, 30, True, 10)
        r["workload"] = ("BASELINE configs[4]: random-policy rollout, 1,048,576 envs per GPU x %d GPU(s) (%d envs), step+mask+obs, episode "
                         "statistics all-reduced over NCCL" % (world, world << 20))
        out["config5"] = r
        torch.cuda.empty_cache()
    except Exception as ex:  # never lose the headline line to a sub-result
        out.setdefault("config4", {"error": repr(ex)})

    if world == 1:
        try:
            out["config3"] = ppo_case(dev, lib)
        except Exception as ex:
            out["config3"] = {"error": repr(ex)}
    return out


def ppo_case(dev, lib, n=262144, steps=6):
    """BASELINE configs[2]: rollout collection with the reference's ActorCritic MLP in the loop (ppo_splendor.py:202-297,
    DualStepNativeWrapper turns), env + masked sampling on the device.  agent-steps/s and the env kernels' share of the time."""
    import ctypes as C

    import torch

    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200 import _lib as L
    from splendor_gym_b200.scripts import ppo_rollout as P

    res = {"workload": "BASELINE configs[2]: PPO self-play rollout, ActorCritic MLP (297-256-256-{45,1}, tanh) in the loop, %d envs on 1 GPU; "
                       "one agent-step = DualStepNativeWrapper.dual_step (agent move + opponent move by the same network)" % n,
           "unit": "agent-steps/s", "envs": n, "dual_steps_timed": steps}
    for name, fmt, dt in (("f16", "f16", torch.float16), ("fp32", "int32", torch.float32)):
        torch.manual_seed(0)
        net = P.ActorCritic().to(dev).to(dt).eval()
        if fmt == "f16":
            net.actor, net.critic = P.pad_head(P.pad_first_layer(net.actor)), P.pad_first_layer(net.critic)
        env = SplendorVecEnv(n, device=dev, seed=42, shuffle="philox", autoreset=True, obs_format=fmt)
        env.reset()
        buf = P.collect(env, net, steps, dtype=dt)
        torch.cuda.synchronize()
        L.check(lib.spl_timing_enable(1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        P.collect(env, net, steps, buffers=buf, dtype=dt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tot, cnt = C.c_double(), C.c_int64()
        L.check(lib.spl_timing_read(C.byref(tot), C.byref(cnt)))
        L.check(lib.spl_timing_enable(0))
        res[name] = {"value": n * steps / (ms * 1e-3), "ms_per_dual_step": ms / steps, "env_step_kernel_ms_per_dual_step": tot.value / steps,
                     "env_step_kernel_share": tot.value / ms if ms > 0 else None, "policy_dtype": str(dt).replace("torch.", ""),
                     "obs_format": fmt}
        env.close()
        del env, buf, net
        torch.cuda.empty_cache()
    res["value"] = res["f16"]["value"]
    return res


# ------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200 import _lib as L

    from splendor_gym_b200.distributed import rank_world

    rank, world, local = rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    N, T, K, W = args.envs, args.rollout, args.steps, max(args.warmup, 3)
    # cap the rollout buffer at ~40 GB
    per_step_bytes = N * (1188 + 45 + 4 + 1 + 4)
    T = max(1, min(T, int(40e9 // per_step_bytes)))
    write_obs = not args.no_obs

    use_rollout = args.mode == "rollout"
    # shuffle="mt19937" (every deal bit-identical to the reference's initial_state(seed)): the rollout kernel takes the deals
    # from a ring of 8 prefetched ones per env (a game lasts >= 17 moves: 8 cover 128 lock-steps), refilled behind the launch
    env = SplendorVecEnv(N, device=dev, seed=20261018, shuffle=args.shuffle, env_offset=rank * N, autoreset=True,
                         prefetch_deals=(args.deal_slots if use_rollout else True))
    obs_buf = torch.zeros((T, N, 297), dtype=torch.int32, device=dev) if write_obs else None
    mask_buf = torch.zeros((T, N, 45), dtype=torch.int8, device=dev)
    rew_buf = torch.zeros((T, N), dtype=torch.float32, device=dev)
    term_buf = torch.zeros((T, N), dtype=torch.uint8, device=dev)
    act_buf = torch.zeros((T + 1, N), dtype=torch.int32, device=dev)
    env.t_base = torch.zeros(1, dtype=torch.int64, device=dev)
    env.reset()
    env.sample_random_actions(out=act_buf[0])

    def segment_lockstep():
        """T lock-steps, one spl_step launch each; the actions for step t+1 are sampled by step t's kernel."""
        for t in range(T):
            env._t = t
            env.step(act_buf[t], out_obs=(obs_buf[t] if write_obs else None), out_mask=mask_buf[t], out_reward=rew_buf[t],
                     out_terminated=term_buf[t], out_next_action=act_buf[t + 1], write_obs=write_obs)
        act_buf[0].copy_(act_buf[T])
        env.t_base += T

    def segment_rollout():
        """The same T lock-steps in ONE launch of the persistent rollout kernel (state stays in registers)."""
        env._t = 0
        env.rollout_random(T, act_buf[0], obs=obs_buf, mask=mask_buf, reward=rew_buf, terminated=term_buf, next_actions=act_buf)
        act_buf[0].copy_(act_buf[T])
        env.t_base += T

    segment = segment_rollout if use_rollout else segment_lockstep
    launches0 = lib.spl_launch_count()
    segment()  # eager once (also validates arguments)
    torch.cuda.synchronize()
    launches_per_segment = lib.spl_launch_count() - launches0

    def make_graph(fn):
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            fn()
        torch.cuda.current_stream().wait_stream(s_)
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            fn()
        return g_

    graph = None
    if not use_rollout and not args.no_graph:
        graph = make_graph(segment_lockstep)
    run = graph.replay if graph is not None else segment

    import ctypes as C

    peak, peak_src = peaks()
    stats_host = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(W):
        run()
    torch.cuda.synchronize()
    live_timing = graph is None  # events around the dominant kernel, recorded on the launching stream in the timed region
    if live_timing:
        L.check(lib.spl_timing_enable(1))  # (creates the event pool: before the barrier, like everything slow and rank-specific)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = torch.cuda.Stream() if world > 1 else None
    if world > 1:
        # one warm all-reduce on the side stream, then the barrier: the ranks enter the timed region together
        with torch.cuda.stream(side):
            dist.all_reduce(stats_host)
        torch.cuda.synchronize()
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k_ in range(K):
        run()
        if world > 1 and ((k_ + 1) % max(1, args.stats_every) == 0 or k_ == K - 1):
            # episode statistics are the only cross-GPU traffic: one NCCL all-reduce of 8 int64 per segment, issued
            # on a side stream so that it overlaps the next segment instead of serialising with it
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                stats_host.copy_(env.stats)
                dist.all_reduce(stats_host)
    if world > 1:
        torch.cuda.current_stream().wait_stream(side)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot, cnt = C.c_double(), C.c_int64()
    if live_timing:
        L.check(lib.spl_timing_read(C.byref(tot), C.byref(cnt)))
        L.check(lib.spl_timing_enable(0))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    total_steps = N * T * K * world
    value = total_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel
    if not live_timing:  # graph replays cannot carry the library's events: time eager launches of the same kernel
        L.check(lib.spl_timing_enable(1))
        for _ in range(max(1, min(K, 4096 // T))):
            segment()
        L.check(lib.spl_timing_read(C.byref(tot), C.byref(cnt)))
        L.check(lib.spl_timing_enable(0))
    clocks = sampler.stop() if sampler else None
    out_bytes = (1188 if write_obs else 0) + 45 + 4 + 1 + 4  # obs + mask + reward + terminated + action, per env-step
    plan = (C.c_int32 * 6)()
    L.check(lib.spl_rollout_plan(N, T, plan))
    if use_rollout:
        kernel = "spl_rollout_kernel"
        units_per_launch = N * T
        # + per work unit (chunk of lock-steps): packed state read + written (128 B) and the first action read (4 B)
        bytes_per_launch = N * T * out_bytes + N * 132 * int(plan[3])
    else:
        kernel = "spl_step_kernel<true>"
        units_per_launch = N
        bytes_per_launch = N * (out_bytes + 128 + 4 + 1)  # + state read + write, action read, info byte
    k_ms = tot.value / max(1, cnt.value)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    survey_bytes = B_FULL if write_obs else B_MASKSTEP
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": kernel, "kernel_ms": k_ms, "launches_timed": cnt.value, "env_steps_per_launch": units_per_launch,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "algorithmic_bytes_per_env_step": bytes_per_launch / units_per_launch,
                "survey_8d_bytes_per_env_step": survey_bytes,
                "frac_with_survey_8d_bytes": units_per_launch * survey_bytes / (k_ms * 1e-3) / 1e9 / peak,
                "peak_source": peak_src, "timed": "live in the timed region" if live_timing else "eager re-run after the graph-timed region"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                tj = json.load(f)
            key = f"{kernel}:{N}:{T}" + ("" if write_obs else ":noobs") + (":mt19937" if use_rollout and args.shuffle == "mt19937" else "")
            if key in tj:
                roofline["traffic"] = tj[key]["dram_bytes_per_launch"]
                roofline["traffic_source"] = tj[key].get("source")
        except Exception:
            pass

    # ---- secondary: the same segment as T separate spl_step launches (policy-in-the-loop path), CUDA graph
    lockstep = None
    if use_rollout and not args.skip_lockstep:
        g2 = make_graph(segment_lockstep)
        for _ in range(2):
            g2.replay()
        k2 = max(2, min(K, 5))
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(k2):
            g2.replay()
        b1.record()
        torch.cuda.synchronize()
        lms = b0.elapsed_time(b1)
        lockstep = {"value": N * T * k2 / (lms * 1e-3), "unit": UNIT, "what": "per-GPU, one spl_step launch per lock-step (CUDA graph of %d launches)" % T,
                    "us_per_lock_step": 1e3 * lms / (k2 * T)}
        del g2
        # the same with the reference's own decks (shuffle="mt19937": every deal bit-identical to initial_state(seed),
        # BASELINE configs[1] "bit-exact replay"); the rollout kernel needs the native Philox deal, this path does not
        env_mt = SplendorVecEnv(N, device=dev, seed=20261018, shuffle="mt19937", env_offset=rank * N, autoreset=True,
                                prefetch_deals=args.deal_slots)
        env_mt.t_base = env.t_base
        env_mt.reset()
        env_mt.sample_random_actions(out=act_buf[0])

        def segment_lockstep_mt():
            for t in range(T):
                env_mt._t = t
                env_mt.step(act_buf[t], out_obs=(obs_buf[t] if write_obs else None), out_mask=mask_buf[t], out_reward=rew_buf[t],
                            out_terminated=term_buf[t], out_next_action=act_buf[t + 1], write_obs=write_obs)
            act_buf[0].copy_(act_buf[T])
            env_mt.t_base += T

        g3 = make_graph(segment_lockstep_mt)
        for _ in range(2):
            g3.replay()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(k2):
            g3.replay()
        b1.record()
        torch.cuda.synchronize()
        lms = b0.elapsed_time(b1)
        lockstep["bit_exact_decks"] = {"value": N * T * k2 / (lms * 1e-3), "unit": UNIT, "us_per_lock_step": 1e3 * lms / (k2 * T),
                                       "what": "same, shuffle=mt19937: decks bit-identical to the reference's initial_state(seed); "
                                               "prefetched next-episode deals (spl_envs_t.spare)"}
        del g3, env_mt
        if args.shuffle == "philox":
            # ... and the rollout kernel itself with the reference's decks: a ring of 8 prefetched MT19937 deals per env,
            # refilled by the batch dealer behind every launch (both inside the timed region)
            env_mt = SplendorVecEnv(N, device=dev, seed=20261018, shuffle="mt19937", env_offset=rank * N, autoreset=True,
                                    prefetch_deals=args.deal_slots)
            env_mt.t_base = env.t_base
            env_mt.reset()
            env_mt.sample_random_actions(out=act_buf[0])

            def segment_rollout_mt():
                env_mt._t = 0
                env_mt.rollout_random(T, act_buf[0], obs=obs_buf, mask=mask_buf, reward=rew_buf, terminated=term_buf, next_actions=act_buf)
                act_buf[0].copy_(act_buf[T])
                env_mt.t_base += T

            def timed(fn, reps):
                for _ in range(3):
                    fn()
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                for _ in range(reps):
                    fn()
                b1.record()
                torch.cuda.synchronize()
                return b0.elapsed_time(b1)

            k3 = max(3, min(K, 20))
            rms = timed(segment_rollout_mt, k3)
            lockstep["rollout_bit_exact_decks"] = {
                "value": N * T * k3 / (rms * 1e-3), "unit": UNIT, "ms_per_segment": rms / k3,
                "what": "per-GPU, the rollout kernel (1 launch per %d lock-steps) with shuffle=mt19937: every deal bit-identical to the "
                        "reference's initial_state(seed), taken from %d prefetched deals per env; the refill (spl_spare_deal_kernel) "
                        "is inside the timed region" % (T, args.deal_slots)}
            del env_mt
        env.sample_random_actions(out=act_buf[0])  # back to the Philox env's own action stream

    # ---- end-to-end through the public API with HOST buffers: SplendorVecEnv.step_host (C ABI spl_host_step).
    # Every lock-step: actions from host memory -> device, step kernel, results -> host memory as the reference-typed
    # arrays (int32 obs, int8 mask, float reward, bool terminated) -- all inside the timed region.
    e2e = None
    e2e_u8 = None
    e2e_light = None
    e2e_plain = None
    if not args.skip_e2e:
        import numpy as np

        n_e2e = max(20, min(200, 4 * T))
        h_act = act_buf[0].cpu().numpy().copy()

        def time_host_loop(step_fn, reps):
            for _ in range(5):
                step_fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0_ = time.perf_counter()
            for _ in range(reps):
                step_fn()
            torch.cuda.synchronize()
            el = (time.perf_counter() - t0_) * 1e3  # host-blocking API: wall clock around the loop == device + host work
            if world > 1:
                t = torch.tensor([el], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                el = float(t.item())
            return el

        def host_step(dtype):
            def f():
                _o, _r, _t, _tr, info = env.step_host(h_act, obs_dtype=dtype, sample_next=True)
                np.copyto(h_act, info["next_action"].numpy())  # the host-side "policy": actions come back from host memory
            return f

        # the node's ceiling for this path: streaming-store rate of the worker pool (uint8 -> int32 widening with
        # non-temporal stores, the loop the path itself runs), measured in-process at the thread count actually used;
        # with several ranks per node all ranks measure at the same time (they share the memory system)
        host_threads = int(lib.spl_host_set_threads(0))
        out_bytes = N * (1188 + 45 + 4 + 1 + 1 + 4)
        if world > 1:
            dist.barrier()
        store_gbs = float(lib.spl_host_store_rate(max(1 << 20, out_bytes // host_threads), 5, 1))
        fill_gbs = float(lib.spl_host_store_rate(max(1 << 20, out_bytes // host_threads), 5, 0))
        if world > 1:
            t = torch.tensor([store_gbs, fill_gbs], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            store_gbs, fill_gbs = float(t[0]), float(t[1])

        launches_e2e0 = lib.spl_launch_count()
        ems = time_host_loop(host_step(torch.int32), n_e2e)
        launches_e2e = (lib.spl_launch_count() - launches_e2e0) // (n_e2e + 5)
        hs = env.host_stats()
        share = hs["gpu_written_share"]
        d2h = int(N * ((1.0 - share) * (148.5 + 17 + 16) + share * 1243) + 4 * ((N + 63) // 64))
        written_gbs = out_bytes * world / (1e3 * ems / n_e2e) * 1e-3
        e2e = {"value": N * n_e2e * world / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * N,
               "d2h_bytes_per_step": d2h, "lock_steps": n_e2e, "us_per_lock_step": 1e3 * ems / n_e2e,
               "host_bytes_written_per_step": out_bytes, "host_threads": host_threads,
               "host_written_gbs": written_gbs, "host_store_gbs": store_gbs, "host_fill_gbs": fill_gbs,
               "frac_of_host_ceiling": written_gbs / store_gbs if store_gbs > 0 else None,
               # the memory system also carries the staging bytes (written by the link, read back by a worker) of the share the
               # host widens: results + 2 x 181.5 B per such env-step -- the path's whole host-memory traffic against the same ceiling
               "host_traffic_bytes_per_step": int(out_bytes + 2 * 181.5 * N * (1.0 - share)),
               "frac_of_host_ceiling_all_traffic": (written_gbs * (out_bytes + 2 * 181.5 * N * (1.0 - share)) / out_bytes) / store_gbs if store_gbs > 0 else None,
               "host_ceiling": "spl_host_store_rate: the worker pool widening uint8 -> int32 with non-temporal stores into %d MB (mode 1; "
                               "host_fill_gbs = plain streaming fill), sustained over >= 40 ms, same threads and pinning as the path, summed over the "
                               "ranks of the node measuring concurrently" % (out_bytes >> 20),
               "gpu_written_share": share, "gpu_writable_results": bool(hs["gpu_writable"]),
               "last_call_us": {k: hs[k] for k in ("call_us", "enqueued_us", "first_group_us", "workers_done_us", "gpu_share_done_us")},
               "kernels_per_lock_step": int(launches_e2e),
               "api": "SplendorVecEnv.step_host(actions) = C ABI spl_host_step: host int32 actions in; host int32 obs [N,297], int8 mask "
                      "[N,45], float reward, bool terminated, info bits, next actions out. Step kernel (compact outputs in HBM) + push kernel: "
                      "64-env groups are stored over PCIe into pinned host memory, nibble-packed (181.5 B per env) for the share that pinned "
                      "host threads widen with non-temporal stores, already widened (1,243 B per env) for the gpu_written_share; the split "
                      "follows the two finish times"}
        ums = time_host_loop(host_step(torch.uint8), n_e2e)
        share8 = env.host_stats()["gpu_written_share"]
        e2e_u8 = {"value": N * n_e2e * world / (ums * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * N,
                  "d2h_bytes_per_step": int(N * ((1.0 - share8) * 181.5 + share8 * 352)), "gpu_written_share": share8,
                  "what": "same call with obs_dtype=uint8: the observation stays bytes on the host (same values; the policy casts to float anyway)"}

        # reference-typed arrays copied as they are (what round 1 first measured): PCIe-bound at 1,243 B per env-step
        p_obs = torch.zeros((N, 297), dtype=torch.int32).pin_memory()
        p_mask = torch.zeros((N, 45), dtype=torch.int8).pin_memory()
        p_rew = torch.zeros(N, dtype=torch.float32).pin_memory()
        p_term = torch.zeros(N, dtype=torch.uint8).pin_memory()
        p_next = torch.zeros(N, dtype=torch.int32).pin_memory()
        p_act = torch.from_numpy(h_act.copy()).pin_memory()
        d_act = torch.zeros(N, dtype=torch.int32, device=dev)

        def plain_step():
            d_act.copy_(p_act, non_blocking=True)
            obs, rew, term, _, info = env.step(d_act, sample_next=True)
            p_obs.copy_(obs, non_blocking=True)
            p_mask.copy_(env.mask, non_blocking=True)
            p_rew.copy_(rew, non_blocking=True)
            p_term.copy_(env._terminated, non_blocking=True)
            p_next.copy_(env.next_action, non_blocking=True)
            torch.cuda.synchronize()
            p_act.copy_(p_next)

        pms = time_host_loop(plain_step, max(20, n_e2e // 4))
        e2e_plain = {"value": N * max(20, n_e2e // 4) * world / (pms * 1e-3), "unit": UNIT, "d2h_bytes_per_step": N * (1188 + 45 + 4 + 1 + 4),
                     "what": "step() + torch copies of the int32/int8/float arrays to pinned memory (no compact form): PCIe-bound"}

        # informational: observation / mask left on the device (a GPU-resident policy consumes them there)
        def light_step():
            d_act.copy_(p_act, non_blocking=True)
            env.step(d_act, sample_next=True)
            p_rew.copy_(env.reward, non_blocking=True)
            p_term.copy_(env._terminated, non_blocking=True)
            p_next.copy_(env.next_action, non_blocking=True)
            torch.cuda.synchronize()
            p_act.copy_(p_next)

        lms_ = time_host_loop(light_step, n_e2e)
        e2e_light = {"value": N * n_e2e * world / (lms_ * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": 9 * N,
                     "what": "obs/mask stay in HBM, actions from and rewards/terminations/next actions to pinned host memory each step"}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, cores, sample = cpu_rollout(min(N, 65536), args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "python_reference": python_reference_record()}

    if world > 1:
        stats_host.copy_(env.stats)
        dist.all_reduce(stats_host)
    else:
        stats_host.copy_(env.stats)
    st = stats_host.cpu().tolist()

    # ---- the other BASELINE configs, bounded (a few seconds each), so that the driver's default run carries them
    extra = {}
    if not args.skip_configs and write_obs and use_rollout:
        env.close()
        del env, obs_buf, mask_buf, rew_buf, term_buf, act_buf
        torch.cuda.empty_cache()
        extra = extra_configs(args, dev, rank, world, lib, peak, peak_src)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(N, T, write_obs),
            "launch": {
                "env_steps_per_step": N * T * world, "shuffle": args.shuffle,
                "mode": "rollout kernel (1 launch per segment)" if use_rollout else "lockstep (1 launch per lock-step)",
                "rollout_plan": {"warps_per_cta": int(plan[0]), "ctas": int(plan[1]), "lock_steps_per_work_unit": int(plan[2]),
                                 "work_units_per_tile_group": int(plan[3]), "tile_groups": int(plan[5]),
                                 "scheduling": "persistent CTAs pull (tile group, step chunk) units from an atomic queue"},
                "cuda_graph": graph is not None, "parallelism": f"env-sharded x{world}, no collective on the step path",
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_u8_obs": e2e_u8, "e2e_plain_copies": e2e_plain, "e2e_device_obs": e2e_light, "lockstep": lockstep,
            "config3": extra.get("config3"), "config4": extra.get("config4"), "config5": extra.get("config5"),
            "gpu_launches": int(launches_per_segment * K), "clocks": clocks,
            "episode_stats": dict(zip(L.STAT_NAMES, st)),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
