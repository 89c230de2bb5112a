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
