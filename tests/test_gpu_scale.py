"""Parity volume: random-policy games replayed bit-exactly against the oracle -- observations, masks, rewards,
terminations, info bits after EVERY step and the full state periodically -- in MT19937 mode (decks identical to the
reference's initial_state).  Default: the north-star volume of >= 1e6 finished games (~8e7 env-steps; about half a minute on the GPU box, result
logged in profiles/r01_parity_volume.log); SPL_SCALE_GAMES overrides the target."""
import os
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_parity_volume(oracle):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from splendor_gym_b200 import SplendorVecEnv

    target = int(os.environ.get("SPL_SCALE_GAMES", "1000000"))
    n = 65536
    env = SplendorVecEnv(n, seed=777, shuffle="mt19937", autoreset=True)
    ref = oracle.OracleVec(n, seed_base=777)
    obs, info = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(obs.cpu().numpy(), robs)
    actions = env.sample_random_actions().clone()
    h_obs = torch.zeros((n, 297), dtype=torch.int32).pin_memory()
    h_mask = torch.zeros((n, 45), dtype=torch.int8).pin_memory()
    t0 = time.time()
    steps = 0
    while True:
        a = actions.cpu().numpy()
        o, r, te, _, inf = env.step(actions, sample_next=True)
        h_obs.copy_(o, non_blocking=True)
        h_mask.copy_(env.mask, non_blocking=True)
        robs, rrew, rterm, rinfo, rmask = ref.step(a, autoreset=True)  # overlaps with the GPU step
        torch.cuda.synchronize()
        assert np.array_equal(h_obs.numpy(), robs), f"obs mismatch at lock-step {steps}"
        assert np.array_equal(h_mask.numpy(), rmask), f"mask mismatch at lock-step {steps}"
        assert np.array_equal(r.cpu().numpy(), rrew) and np.array_equal(env._terminated.cpu().numpy(), rterm)
        assert np.array_equal(env.info_bits.cpu().numpy(), rinfo)
        actions = env.next_action.clone()
        steps += 1
        if steps % 100 == 0:
            assert np.array_equal(env.export_state().cpu().numpy(), ref.export_rows()), f"state mismatch at lock-step {steps}"
        games = int(ref.stats()[0])
        if games >= target:
            break
    assert np.array_equal(env.export_state().cpu().numpy(), ref.export_rows())
    assert np.array_equal(env.stats.cpu().numpy(), ref.stats())
    st = ref.stats()
    msg = (f"parity volume: {games} games, {steps} lock-steps x {n} envs = {steps * n} env-steps bit-exact "
           f"(p0 {st[1]}, p1 {st[2]}, ties {st[3]}, limit {st[4]}, no-legal {st[5]}) in {time.time() - t0:.1f} s")
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_volume.log"), "a") as f:
            f.write(msg + "\n")


@pytest.mark.parametrize("async_refill", [False, True])
def test_parity_volume_rollout_kernel(oracle, async_refill):
    """The same volume through spl_rollout_random with the reference's own decks (shuffle='mt19937', ring of 8 prefetched
    deals per env, batch dealer behind every 128-step launch): every observation, mask, reward, termination and info
    byte of every lock-step against the oracle, which deals with CPython's random.Random(seed).shuffle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from splendor_gym_b200 import SplendorVecEnv

    target = int(os.environ.get("SPL_SCALE_GAMES_ROLLOUT", os.environ.get("SPL_SCALE_GAMES", "1000000")))
    if async_refill:  # refills on a side stream while the next launch runs (two launches in flight before the checks)
        target //= 4
    n, T = 65536, 128
    dev = torch.device("cuda")
    env = SplendorVecEnv(n, seed=4321, shuffle="mt19937", autoreset=True, prefetch_deals=16 if async_refill else 8)
    side = torch.cuda.Stream()
    pending = []
    ref = oracle.OracleVec(n, seed_base=4321)
    obs0, _ = env.reset()
    robs, _ = ref.reset()
    assert np.array_equal(obs0.cpu().numpy(), robs)
    obs = torch.zeros((T, n, 297), dtype=torch.int32, device=dev)
    mask = torch.zeros((T, n, 45), dtype=torch.int8, device=dev)
    rew = torch.zeros((T, n), dtype=torch.float32, device=dev)
    term = torch.zeros((T, n), dtype=torch.uint8, device=dev)
    info = torch.zeros((T, n), dtype=torch.uint8, device=dev)
    acts = torch.zeros((T + 1, n), dtype=torch.int32, device=dev)
    acts[0] = env.sample_random_actions()
    t0 = time.time()
    steps = 0
    games = 0
    while games < target:
        if async_refill:
            # the deals this launch takes are replaced on the side stream WHILE the host checks run and -- because the
            # wait is for the refill before last -- while the next launch runs
            main = torch.cuda.current_stream()
            if len(pending) >= 2:
                main.wait_event(pending[-2])
            env.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts, info=info, refill=False)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                env.refill_deals()
                ev = torch.cuda.Event()
                ev.record(side)
            pending.append(ev)
        else:
            env.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts, info=info)
        h_acts = acts.cpu().numpy()
        h_small = [x.cpu().numpy() for x in (rew, term, info)]
        for t in range(T):
            robs, rrew, rterm, rinfo, rmask = ref.step(h_acts[t], autoreset=True)
            assert torch.equal(obs[t], torch.from_numpy(robs).to(dev)), f"obs mismatch at lock-step {steps}"
            assert torch.equal(mask[t], torch.from_numpy(rmask).to(dev)), f"mask mismatch at lock-step {steps}"
            assert np.array_equal(h_small[0][t], rrew) and np.array_equal(h_small[1][t], rterm) and np.array_equal(h_small[2][t], rinfo), \
                f"reward / terminated / info mismatch at lock-step {steps}"
            steps += 1
        assert np.array_equal(env.export_state().cpu().numpy(), ref.export_rows()), f"state mismatch after lock-step {steps}"
        acts[0].copy_(acts[T])
        games = int(ref.stats()[0])
    assert np.array_equal(env.stats.cpu().numpy(), ref.stats())
    st = ref.stats()
    torch.cuda.synchronize()
    msg = (f"parity volume (rollout kernel, MT19937 decks{', refills on a side stream' if async_refill else ''}): {games} games, {steps} lock-steps x {n} envs = {steps * n} env-steps "
           f"bit-exact (p0 {st[1]}, p1 {st[2]}, ties {st[3]}, limit {st[4]}, no-legal {st[5]}) in {time.time() - t0:.1f} s")
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_volume.log"), "a") as f:
            f.write(msg + "\n")
