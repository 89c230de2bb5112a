"""Throughput of the UNMODIFIED Python reference engine (test / measurement infrastructure, runs only where
/root/reference exists, i.e. in the build container -- never on the GPU box).

BASELINE.md section 4 / north_star name this baseline: `SplendorEnv` stepped with uniformly random legal actions
(splendor_gym/scripts/random_rollout.py:13-30) under a vector env on the host cores --
  * "sync":  one process stepping `--envs` SplendorEnv instances in a Python loop with same-step auto-reset, which is
             what gymnasium.vector.SyncVectorEnv does (ppo_splendor.py:151); gymnasium itself is not installed in this
             image, so the loop is written out (oracle/pyref.py supplies the gymnasium.Env stand-in the engine imports);
  * "async": one such loop per core in separate processes = an upper bound for gymnasium.vector.AsyncVectorEnv
             (which adds a pipe round trip per env per step on top).
Each leg runs for >= `--seconds`.  Output: tests/golden/pyref_throughput.json (read by bench.py, labelled "other box").

  python oracle/time_pyref.py [--seconds 20] [--envs 16] [--procs N]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sync_loop(envs: int, seconds: float, seed: int = 0):
    import numpy as np

    from oracle import pyref

    ref = pyref.load()
    E = [ref.env.SplendorEnv(num_players=2) for _ in range(envs)]
    masks = []
    for i, e in enumerate(E):
        _, info = e.reset(seed=seed * 100003 + i)
        masks.append(info["action_mask"])
    rng = np.random.RandomState(seed)
    steps = episodes = 0
    t0 = time.perf_counter()
    while True:
        for i, e in enumerate(E):
            legal = np.flatnonzero(masks[i])
            a = int(legal[rng.randint(len(legal))]) if len(legal) else 0
            _, _, term, trunc, info = e.step(a)
            if term or trunc:  # same-step auto-reset, as the vector env does
                episodes += 1
                _, info = e.reset()
            masks[i] = info["action_mask"]
        steps += envs
        el = time.perf_counter() - t0
        if el >= seconds:
            return steps, episodes, el


def _worker(args):
    envs, seconds, seed = args
    return sync_loop(envs, seconds, seed)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--envs", type=int, default=16, help="envs per process (ppo_splendor.py --num-envs default)")
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "pyref_throughput.json"))
    a = ap.parse_args()
    s_steps, s_eps, s_el = sync_loop(a.envs, a.seconds)
    with mp.get_context("fork").Pool(a.procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_worker, [(a.envs, a.seconds, 1 + k) for k in range(a.procs)])
        wall = time.perf_counter() - t0
    a_steps = sum(r[0] for r in res)
    rec = {
        "what": "unmodified reference engine (splendor_gym.SplendorEnv, random legal actions, same-step auto-reset), Python loops",
        "unit": "env-steps/s",
        "sync": {"value": s_steps / s_el, "processes": 1, "envs": a.envs, "seconds": s_el, "env_steps": s_steps, "episodes": s_eps},
        "async": {"value": a_steps / max(r[2] for r in res), "processes": a.procs, "envs_per_process": a.envs, "seconds": max(r[2] for r in res),
                  "wall_seconds_incl_start": wall, "env_steps": a_steps, "episodes": sum(r[1] for r in res),
                  "note": "one SyncVectorEnv-style loop per core, no IPC: an upper bound for gymnasium.vector.AsyncVectorEnv"},
        "host": {"cpu": cpu_model(), "cores_used": a.procs, "python": platform.python_version(), "machine": platform.node() and "build container"},
        "generated_by": "oracle/time_pyref.py",
    }
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
        f.write("\n")
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
