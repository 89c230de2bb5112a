"""N > 1 plumbing on CPU with two gloo ranks: contiguous env sharding by global index, rank-independent seed
schedule (a 2-rank run reproduces the 1-rank run env by env), episode-statistics all-reduce, and bench.py's
reference arm under a multi-rank launch (rank 0 prints, the others exit 0 without work).  The env work itself
is done by the oracle here -- the GPU kernels use the same (seed_base, global env id, episode) schedule and
are checked against the oracle in tests/test_gpu_parity.py."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from splendor_gym_b200.distributed import rank_world, shard, all_reduce_stats
from oracle import oracle as O
rank, world, local = rank_world()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
total, steps = 96, 120
off, n = shard(total, rank, world)
v = O.OracleVec(n, seed_base=17, env_offset=off)
v.reset()
for t in range(steps):
    v.step(v.random_actions(0xB200, t), autoreset=True)
stats = all_reduce_stats(torch.from_numpy(v.stats().copy()))
rows = v.export_rows()
gathered = [None] * world
dist.all_gather_object(gathered, (off, rows))
if rank == 0:
    rows_all = np.concatenate([r for _, r in sorted(gathered, key=lambda x: x[0])])
    np.save(%(out)r, rows_all)
    print(json.dumps({"stats": stats.tolist(), "shards": [(o, len(r)) for o, r in gathered]}))
dist.destroy_process_group()
"""


def test_shard_helper():
    from splendor_gym_b200.distributed import shard

    for total, world in ((96, 2), (100, 8), (7, 3), (1 << 23, 8)):
        parts = [shard(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == total
        for (o1, n1), (o2, _) in zip(parts, parts[1:]):
            assert o1 + n1 == o2
    assert shard(1 << 23, 3, 8) == (3 << 20, 1 << 20)


def test_two_gloo_ranks_reproduce_single_rank(tmp_path, oracle):
    out = str(tmp_path / "rows.npy")
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port, "out": out})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert sorted(res["shards"]) == [[0, 48], [48, 48]]
    # single process over all 96 envs
    v = oracle.OracleVec(96, seed_base=17, env_offset=0)
    v.reset()
    for t in range(120):
        v.step(v.random_actions(0xB200, t), autoreset=True)
    assert np.array_equal(np.load(out), v.export_rows())
    assert res["stats"] == v.stats().tolist() and res["stats"][0] > 0


def test_bench_reference_arm_multi_rank():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints one JSON line, rank 1 exits 0 silently."""
    outs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), OMP_NUM_THREADS="4")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                            "--envs", "2048"], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip())
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["n_gpus"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and outs[1] == ""


HOST_WORKER = r"""
import os, sys, json, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np
from splendor_gym_b200 import _lib
lib = _lib.load()
threads = lib.spl_host_set_threads(0)          # default pool of this rank
n = 5000 + 64 * int(os.environ["LOCAL_RANK"])
rng = np.random.default_rng(n)
obs8 = rng.integers(0, 256, size=(n, 297), dtype=np.uint8)
mbits = rng.integers(0, 1 << 45, size=n, dtype=np.uint64)
side = np.zeros((n, 4), np.uint32)
side[:, 0] = (mbits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
side[:, 1] = (mbits >> np.uint64(32)).astype(np.uint32) | (np.uint32(1) << 16)
obs, mask, rew = np.zeros((n, 297), np.int32), np.zeros((n, 45), np.int8), np.zeros(n, np.float32)
io = _lib.SplHostIO(obs=obs.ctypes.data, mask=mask.ctypes.data, reward=rew.ctypes.data)
for _ in range(20):
    assert lib.spl_host_expand(obs8.ctypes.data, side.ctypes.data, n, C.byref(io)) == 0
want = ((mbits[:, None] >> np.arange(45, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8)
ok = bool(np.array_equal(obs, obs8.astype(np.int32)) and np.array_equal(mask, want) and np.all(rew == 1.0))
rate = lib.spl_host_store_rate(1 << 20, 2, 1)
print(json.dumps({"threads": threads, "ok": ok, "rate_positive": rate > 0, "cores": len(os.sched_getaffinity(0))}))
"""


def test_host_pool_under_two_local_ranks(tmp_path):
    """The host half of the host-buffer path (worker pool, per-rank core slices, widening) with two ranks of one node running at
    the same time: each rank sizes its pool from its share of the cores (cores / LOCAL_WORLD_SIZE minus two, at least 3/4 of them), pins its workers
    inside its own slice and widens correctly; the caller's affinity mask is left as it was."""
    script = tmp_path / "host_worker.py"
    script.write_text(HOST_WORKER % {"root": ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), LOCAL_WORLD_SIZE="2")
        env.pop("SPL_HOST_THREADS", None)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    cores = len(os.sched_getaffinity(0))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
        rec = json.loads(out.strip().splitlines()[-1])
        assert rec["ok"] and rec["rate_positive"]
        per = max(1, cores // 2)
        assert rec["threads"] == max(1, min(32, max(per - 2, (per * 3 + 3) // 4)))
        assert rec["cores"] == cores  # the calling thread's mask is restored after the rank-local first touch
