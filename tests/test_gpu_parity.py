"""GPU parity tests: the CUDA path (through the C ABI, via SplendorVecEnv) against the oracle and the
golden vectors generated from the reference.  Bit-exact: observations, masks, rewards, terminations,
info bits and the full exported state after every step."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def VecEnv():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from splendor_gym_b200 import SplendorVecEnv

    return SplendorVecEnv


def assert_step_equal(env, out, ref_out, t, check_state=None):
    obs, rew, term, trunc, info = out
    robs, rrew, rterm, rinfo, rmask = ref_out
    o, m = _np(obs), _np(info["action_mask"])
    bad = np.flatnonzero((o != robs).any(axis=1))
    assert bad.size == 0, f"step {t}: obs mismatch in envs {bad[:8]} cols {np.flatnonzero(o[bad[0]] != robs[bad[0]])[:12]}"
    bad = np.flatnonzero((m != rmask).any(axis=1))
    assert bad.size == 0, f"step {t}: mask mismatch in envs {bad[:8]}"
    assert np.array_equal(_np(rew), rrew), f"step {t}: reward"
    assert np.array_equal(_np(term).astype(np.uint8), rterm), f"step {t}: terminated"
    assert np.array_equal(_np(info["info_bits"]), rinfo), f"step {t}: info bits"
    assert not _np(trunc).any()
    if check_state is not None:
        rows = _np(env.export_state())
        bad = np.flatnonzero((rows != check_state).any(axis=1))
        assert bad.size == 0, f"step {t}: state mismatch env {bad[:4]} fields {np.flatnonzero(rows[bad[0]] != check_state[bad[0]])[:12]}"


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000])
def test_reset_mt19937_bit_exact(VecEnv, oracle, n):
    """spl_reset in MT mode == initial_state(seed) (engine/state.py:181-211), ragged tile sizes included."""
    env = VecEnv(n, seed=7, shuffle="mt19937", env_offset=5)
    obs, info = env.reset()
    ref = oracle.OracleVec(n, seed_base=7, env_offset=5)
    robs, rmask = ref.reset()
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(obs), robs) and np.array_equal(_np(info["action_mask"]), rmask)


def test_reset_explicit_engine_seeds(VecEnv, oracle):
    seeds = [0, 42, 123456789, 2**31 - 2, 1826701614, 191664963, 33158374, 2**32 + 5, 2**40 + 7]
    env = VecEnv(len(seeds), shuffle="mt19937")
    env.reset(seeds=torch.tensor(seeds, dtype=torch.int64))
    rows = _np(env.export_state())
    for i, s in enumerate(seeds):
        assert np.array_equal(rows[i], oracle.initial_row(s)), s
    g = {r["seed"]: r for r in load_golden("initial_states.json")}
    for i, s in enumerate(seeds):
        if s in g:
            assert rows[i].tolist() == g[s]["row"]
            assert _np(env.obs)[i].tolist() == oracle.encode_observation(np.array(g[s]["row"], np.int32)).tolist()


@pytest.mark.parametrize("n,steps,illegal_rate", [(33, 260, 0.0), (4096, 300, 0.0), (2048, 200, 0.05)])
def test_lockstep_rollout_bit_exact(VecEnv, oracle, n, steps, illegal_rate):
    """Random-policy lock-step rollout with same-step auto-reset; every output of every step and the
    full state are compared with the oracle (which replays the reference's rules on the same actions)."""
    env = VecEnv(n, seed=2024, shuffle="mt19937", autoreset=True)
    ref = oracle.OracleVec(n, seed_base=2024)
    obs, info = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(_np(obs), robs)
    rng = np.random.RandomState(n)
    actions = env.sample_random_actions().clone()
    assert np.array_equal(_np(actions), ref.random_actions(env.action_key, 0))
    for t in range(steps):
        a = _np(actions).copy()
        if illegal_rate > 0:
            flip = rng.rand(n) < illegal_rate
            a[flip] = rng.randint(-3, 50, size=int(flip.sum()))
        out = env.step(torch.from_numpy(a).cuda(), sample_next=True)
        ref_out = ref.step(a, autoreset=True)
        assert_step_equal(env, out, ref_out, t, check_state=ref.export_rows() if t % 25 == 0 or t == steps - 1 else None)
        actions = env.next_action.clone()
        # fused sampler == standalone sampler == oracle's stream
        assert np.array_equal(_np(actions), ref.random_actions(env.action_key, t + 1)), f"step {t}: sampled actions"
    assert np.array_equal(_np(env.stats), ref.stats())
    assert np.array_equal(_np(env.episode).astype(np.uint32), ref.episodes())
    assert ref.stats()[0] > 0


def test_edge_cases_from_reference_tests(VecEnv, oracle):
    """Hand-built states mirroring the reference's own tests (tests/golden/edge_cases.json), injected with
    import_state exactly as those tests mutate env.state."""
    cases = load_golden("edge_cases.json")
    n = len(cases)
    env = VecEnv(n, shuffle="mt19937", autoreset=False)
    env.import_state(torch.tensor([c["row_in"] for c in cases], dtype=torch.int32))
    assert np.array_equal(_np(env.export_state()), np.array([c["row_in"] for c in cases], np.int32))
    obs0, mask0 = env.observe()
    for i, c in enumerate(cases):
        if c.get("raises") != "RuntimeError":
            assert _np(mask0)[i].tolist() == c["mask_in"], c["name"]
    obs, rew, term, trunc, info = env.step(torch.tensor([c["action"] for c in cases], dtype=torch.int32))
    rows = _np(env.export_state())
    o, m, r, te, ib = _np(obs), _np(info["action_mask"]), _np(rew), _np(term), _np(info["info_bits"])
    for i, c in enumerate(cases):
        if "raises" in c:
            assert ib[i] & 64 and rows[i].tolist() == c["row_in"], c["name"]
            assert bool(ib[i] & 8) == (c["raises"] == "RuntimeError")
            continue
        assert rows[i].tolist() == c["row_out"], c["name"]
        assert o[i].tolist() == c["obs"], c["name"]
        assert m[i].tolist() == c["mask"], c["name"]
        assert r[i] == pytest.approx(c["reward"]) and bool(te[i]) == c["terminated"] and int(ib[i]) == c["info"], c["name"]
    r0, r1, present = env.final_rewards(info["info_bits"])
    for i, c in enumerate(cases):
        if c.get("final_rewards") is not None:
            assert bool(present[i]) and [float(r0[i]), float(r1[i])] == pytest.approx(c["final_rewards"]), c["name"]
        elif "raises" not in c:
            assert not bool(present[i]), c["name"]


def test_golden_games_replay(VecEnv):
    """Full reference games (tests/golden/games.json) replayed on the GPU, one env per game, lock-step."""
    from test_oracle_golden import digest

    games = load_golden("games.json")
    n = len(games)
    env = VecEnv(n, shuffle="mt19937", autoreset=False)
    env.reset(seeds=torch.tensor([g["seed"] for g in games], dtype=torch.int64))
    T = max(len(g["actions"]) for g in games)
    done = np.zeros(n, bool)
    for t in range(T):
        a = np.array([g["actions"][t] if t < len(g["actions"]) else 0 for g in games], np.int32)
        active = np.array([t < len(g["actions"]) for g in games])
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda(), active=torch.from_numpy(active).cuda())
        rows = _np(env.export_state())
        o, m, r, te, ib = _np(obs), _np(info["action_mask"]), _np(rew), _np(term), _np(info["info_bits"])
        for i, g in enumerate(games):
            if active[i]:
                assert digest(o[i], m[i], rows[i], r[i], te[i], int(ib[i])) == g["digests"][t], (g["seed"], g["policy"], t)
    rows = _np(env.export_state())
    for i, g in enumerate(games):
        assert int(rows[i][72]) == g["moves"]
        assert (None if rows[i][74] < 0 else int(rows[i][74])) == g["winner"]


def test_philox_mode_rollout(VecEnv, oracle):
    """Native Philox shuffles are not bit-equal to the reference's MT19937 decks, so the oracle is fed the
    GPU's dealt state (replay of deck permutations) at every reset; everything else must be bit-exact."""
    n, steps = 2048, 220
    env = VecEnv(n, seed=99, shuffle="philox", autoreset=True)
    obs, info = env.reset()
    rows = _np(env.export_state())
    # a valid deal: every card exactly once across board + decks, 3 distinct nobles, fresh counters
    for i in range(0, n, 97):
        r = rows[i]
        cards = list(r[52:64]) + list(r[76:76 + r[64]]) + list(r[116:116 + r[65]]) + list(r[146:146 + r[66]])
        assert sorted(cards) == list(range(90)) and (r[64], r[65], r[66]) == (36, 26, 16)
        assert all(0 <= c < 40 for c in r[52:56]) and all(40 <= c < 70 for c in r[56:60]) and all(70 <= c < 90 for c in r[60:64])
        assert len(set(r[67:70])) == 3 and all(0 <= x < 10 for x in r[67:70])
    assert len({tuple(r[52:64]) for r in rows}) > n * 0.99  # decks differ across envs
    ref = oracle.OracleVec(n)
    ref.import_rows(rows)
    robs, rmask = ref.observe()
    assert np.array_equal(_np(obs), robs) and np.array_equal(_np(info["action_mask"]), rmask)
    actions = env.sample_random_actions().clone()
    resets = 0
    for t in range(steps):
        a = _np(actions).copy()
        obs, rew, term, trunc, info = env.step(actions, sample_next=True)
        robs, rrew, rterm, rinfo, rmask = (x.copy() for x in ref.step(a, autoreset=False))
        assert np.array_equal(_np(rew), rrew) and np.array_equal(_np(term).astype(np.uint8), rterm)
        ib = _np(info["info_bits"])
        assert np.array_equal(ib & 0x7F, rinfo)
        was_reset = (ib & 128) != 0
        assert np.array_equal(was_reset, rterm != 0)
        if was_reset.any():
            rows = _np(env.export_state())
            assert (rows[was_reset][:, 72] == 0).all()
            ref.import_rows(rows, which=was_reset.astype(np.uint8))
            o2, m2 = ref.observe()
            robs[was_reset], rmask[was_reset] = o2[was_reset], m2[was_reset]
            resets += int(was_reset.sum())
        assert np.array_equal(_np(obs), robs), f"step {t}"
        assert np.array_equal(_np(info["action_mask"]), rmask), f"step {t}"
        actions = env.next_action.clone()
    assert resets > n
    assert np.array_equal(_np(env.export_state()), ref.export_rows())


def test_philox_deals_are_uniform(VecEnv):
    """First-slot card of each tier and first noble should be uniform over 40/30/20/10 values."""
    n = 65536
    env = VecEnv(n, seed=3, shuffle="philox")
    env.reset()
    rows = _np(env.export_state())
    for col, lo, k in ((52, 0, 40), (56, 40, 30), (60, 70, 20), (67, 0, 10), (76 + 35, 0, 40)):
        counts = np.bincount(rows[:, col] - lo, minlength=k)
        expected = n / k
        chi2 = ((counts - expected) ** 2 / expected).sum()
        assert chi2 < 3.0 * k, (col, chi2)  # dof = k-1; 3x is far beyond the 1e-6 tail


def test_active_mask_and_observe(VecEnv, oracle):
    n = 777
    env = VecEnv(n, seed=5, shuffle="mt19937", autoreset=False)
    env.reset()
    ref = oracle.OracleVec(n, seed_base=5)
    ref.reset()
    rng = np.random.RandomState(0)
    for t in range(60):
        a = ref.random_actions(1, t)
        active = rng.rand(n) < 0.6
        out = env.step(torch.from_numpy(a).cuda(), active=torch.from_numpy(active).cuda())
        robs, rrew, rterm, rinfo, rmask = ref.step(a, active=active.astype(np.uint8), autoreset=False)
        rrew, rterm, rinfo = rrew.copy(), rterm.copy(), rinfo.copy()
        rrew[~active], rterm[~active], rinfo[~active] = 0, 0, 0
        o2, m2 = ref.observe()
        assert_step_equal(env, out, (o2, rrew, rterm, rinfo, m2), t, check_state=ref.export_rows())
    o, m = env.observe()
    o2, m2 = ref.observe()
    assert np.array_equal(_np(o), o2) and np.array_equal(_np(m), m2)


def test_random_action_kernel_matches_oracle_stream(VecEnv, oracle):
    n = 1000
    env = VecEnv(n, seed=11, shuffle="mt19937")
    env.reset()
    ref = oracle.OracleVec(n, seed_base=11)
    ref.reset()
    got = _np(env.sample_random_actions())
    assert np.array_equal(got, ref.random_actions(env.action_key, 0))
    m = _np(env.mask)
    assert (m[np.arange(n), got] == 1).all()
    zero = torch.zeros((n, 45), dtype=torch.int8, device="cuda")
    assert (_np(env.sample_random_actions(mask=zero)) == 0).all()


def test_dual_step_matches_two_reference_steps(VecEnv, oracle):
    """DualStepNativeWrapper.dual_step = env.step(agent) -> opponent policy -> env.step(opponent)
    (wrappers/dual_step_native.py:90-193) with the reward bookkeeping of :132-167."""
    n = 1024
    env = VecEnv(n, seed=77, shuffle="mt19937", autoreset=True)
    ref = oracle.OracleVec(n, seed_base=77)
    env.reset()
    ref.reset()
    t = [0]

    def opp_policy(obs, mask):
        return env.sample_random_actions(mask)

    dones = 0
    for it in range(150):
        a = env.sample_random_actions().clone()
        a_np = _np(a).copy()
        agent_obs, agent_r, opp_obs, opp_r, done, info = env.dual_step(a, opp_policy)
        # oracle: two single steps with the same actions
        _, r1, t1, i1, _ = (x.copy() for x in ref.step(a_np, autoreset=True))
        active = (t1 == 0) & ((i1 & 65) == 0)
        opp_a = ref.random_actions(env.action_key, env._t - 1)
        robs, r2, t2, i2, rmask = (x.copy() for x in ref.step(opp_a, active=active.astype(np.uint8), autoreset=True))
        robs, rmask = ref.observe()
        w2 = ((i2 >> 4) & 3).astype(int) - 1
        fr0 = np.where(w2 < 0, np.where(i2 & 4, -0.1, 0.0), np.where(w2 == 0, 1.0, -1.0)).astype(np.float32)
        fr0 = np.where((t2 != 0) & ((i2 & 2) == 0), fr0, 0.0).astype(np.float32)
        w1 = ((i1 >> 4) & 3).astype(int) - 1
        fr1 = np.where(w1 < 0, np.where(i1 & 4, -0.1, 0.0), np.where(w1 == 1, 1.0, -1.0)).astype(np.float32)
        fr1 = np.where((t1 != 0) & ((i1 & 2) == 0), fr1, 0.0).astype(np.float32)
        want_agent = np.where(t1 != 0, r1, np.where(active, np.where(t2 != 0, fr0, 0.0), r1)).astype(np.float32)
        want_opp = np.where(t1 != 0, fr1, np.where(active, r2, 0.0)).astype(np.float32)
        want_done = (t1 != 0) | (active & (t2 != 0))
        assert np.array_equal(_np(agent_r), want_agent) and np.array_equal(_np(opp_r), want_opp)
        assert np.array_equal(_np(done), want_done)
        assert np.array_equal(_np(agent_obs), robs) and np.array_equal(_np(info["action_mask"]), rmask)
        assert (_np(agent_obs)[:, 294] == 0).all()  # agent (player 0) to move again
        dones += int(want_done.sum())
    assert dones > 500


def test_full_size_invariants(VecEnv):
    """BASELINE config sizes (65,536 and 1,048,576 envs): size-independent properties after a rollout --
    token conservation per colour, hand limit, observation/mask/state consistency through spl_observe."""
    for n in (65536, 1 << 20):
        env = VecEnv(n, seed=1, shuffle="philox", autoreset=True)
        env.reset()
        actions = env.sample_random_actions()
        for t in range(48):
            obs, rew, term, trunc, info = env.step(actions, sample_next=True)
            actions = env.next_action
        obs = env.obs
        total = obs[:, 0:6] + obs[:, 6:12] + obs[:, 19:25]
        want = torch.tensor([4, 4, 4, 4, 4, 5], dtype=torch.int32, device=obs.device)
        assert bool((total == want).all())
        assert int(obs[:, 6:12].sum(dim=1).max()) <= 10 and int(obs[:, 19:25].sum(dim=1).max()) <= 10
        assert int(obs.min()) >= 0 and int(obs[:, 296].max()) == 0  # auto-reset: never left terminal
        o1, m1 = obs.clone(), env.mask.clone()
        o2, m2 = env.observe()
        assert torch.equal(o1, o2) and torch.equal(m1, m2)
        # a non-terminal state may have NO legal move (the next step turns it into a draw, envs/splendor_env.py:55-61);
        # everywhere else the sampled action is legal under the returned mask, and 0 where nothing is legal
        has_move = m1.sum(dim=1) > 0
        assert float(has_move.float().mean()) > 0.99
        picked = m1.gather(1, env.next_action.long().view(-1, 1)).view(-1)
        assert bool((picked[has_move] == 1).all()) and bool((env.next_action[~has_move] == 0).all())
        assert int(env.stats[0]) > 0
        del env
        torch.cuda.empty_cache()


def test_full_size_sampled_parity(VecEnv, oracle):
    """65,536 envs x 96 lock-steps bit-exact against the oracle (every env, every step: obs + mask checksum
    on device, full compare on a strided sample)."""
    n, steps = 65536, 96
    env = VecEnv(n, seed=42, shuffle="mt19937", autoreset=True)
    ref = oracle.OracleVec(n, seed_base=42)
    env.reset()
    ref.reset()
    actions = env.sample_random_actions().clone()
    for t in range(steps):
        a = _np(actions)
        out = env.step(actions, sample_next=True)
        ref_out = ref.step(a, autoreset=True)
        assert_step_equal(env, out, ref_out, t)
        actions = env.next_action.clone()
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(env.stats), ref.stats())


@pytest.mark.parametrize("n", [48, 1000, 8192])
def test_rollout_kernel_equals_chained_steps(VecEnv, oracle, n):
    """spl_rollout_random (T lock-steps in one launch, fused Philox auto-reset) is bit-identical to T chained
    spl_step calls, and both follow the oracle when it is handed the dealt states."""
    T = 150
    a = VecEnv(n, seed=31, shuffle="philox", autoreset=True)
    b = VecEnv(n, seed=31, shuffle="philox", autoreset=True)
    a.reset()
    b.reset()
    assert torch.equal(a.export_state(), b.export_state())
    act0 = a.sample_random_actions().clone()
    obs = torch.zeros((T, n, 297), dtype=torch.int32, device="cuda")
    mask = torch.zeros((T, n, 45), dtype=torch.int8, device="cuda")
    rew = torch.zeros((T, n), dtype=torch.float32, device="cuda")
    term = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
    info = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
    acts = torch.zeros((T + 1, n), dtype=torch.int32, device="cuda")
    acts[0] = act0
    a.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts, info=info)
    actions = act0.clone()
    for t in range(T):
        assert torch.equal(actions, acts[t]), f"step {t}: action chain"
        o, r, te, _, inf = b.step(actions, sample_next=True)
        assert torch.equal(o, obs[t]), f"step {t}: obs"
        assert torch.equal(b.mask, mask[t]), f"step {t}: mask"
        assert torch.equal(r, rew[t]) and torch.equal(b._terminated, term[t]) and torch.equal(b.info_bits, info[t])
        actions = b.next_action.clone()
    assert torch.equal(actions, acts[T])
    assert torch.equal(a.export_state(), b.export_state())
    assert torch.equal(a.stats, b.stats) and torch.equal(a.episode, b.episode)
    assert int(a.stats[0]) > 0


@pytest.mark.parametrize("chunk,wpc,ctas_per_sm", [(1, 1, 0), (3, 4, 0), (8, 4, 1), (16, 1, 2), (64, 4, 0), (5, 1, 20)])
def test_rollout_schedule_independence(VecEnv, monkeypatch, chunk, wpc, ctas_per_sm):
    """The rollout kernel pulls (tile group, step chunk) work units from an atomic queue; its outputs must not depend on
    the chunk length, the CTA shape or the number of CTAs (i.e. on which SM ran which chunk, and in what order)."""
    n, T = 4096 + 48, 70

    def run():
        e = VecEnv(n, seed=77, shuffle="philox", autoreset=True)
        e.reset()
        acts = torch.zeros((T + 1, n), dtype=torch.int32, device="cuda")
        acts[0] = e.sample_random_actions()
        obs = torch.zeros((T, n, 297), dtype=torch.int32, device="cuda")
        mask = torch.zeros((T, n, 45), dtype=torch.int8, device="cuda")
        rew = torch.zeros((T, n), dtype=torch.float32, device="cuda")
        term = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        info = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        e.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts, info=info)
        torch.cuda.synchronize()
        return obs, mask, rew, term, info, acts, e.export_state(), e.stats.clone(), e.episode.clone()

    ref = run()
    monkeypatch.setenv("SPL_ROLLOUT_CHUNK", str(chunk))
    monkeypatch.setenv("SPL_WPC", str(wpc))
    if ctas_per_sm:
        monkeypatch.setenv("SPL_ROLLOUT_CTAS_PER_SM", str(ctas_per_sm))
    out = run()
    for a, b in zip(ref, out):
        assert torch.equal(a, b)
    assert int(ref[7][0]) > 0


@pytest.mark.parametrize("age,prefetch", [(None, True), ("100000", True), ("1", True), (None, False), ("1:130", True), (None, 1), ("16", 1)])
def test_mt19937_prefetched_deals_bit_exact(VecEnv, oracle, monkeypatch, age, prefetch):
    """shuffle='mt19937' with auto-reset takes the prefetched deal of each env's next episode (spl_envs_t.spare) and refills the
    spares in batches.  Whatever the refill cadence -- default, never (every later finish then falls back to the in-line
    reset kernel), every lock-step -- and without spares at all, every output stays bit-identical to the oracle."""
    if age is not None:
        if ":" in age:  # also hand about half of the deals back to the consumer (see test_mt19937_rollout_kernel_bit_exact)
            age, max_out = age.split(":")
            monkeypatch.setenv("SPL_DEAL_MAX_OUTPUTS", max_out)
        monkeypatch.setenv("SPL_SPARE_REFILL_AGE", age)
    n, steps = 2048, 260
    env = VecEnv(n, seed=99, shuffle="mt19937", autoreset=True, prefetch_deals=prefetch)
    assert (env.spare is not None) == bool(prefetch)
    ref = oracle.OracleVec(n, seed_base=99)
    obs, info = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(_np(obs), robs) and np.array_equal(_np(info["action_mask"]), rmask)
    actions = env.sample_random_actions().clone()
    for t in range(steps):
        a = actions.clone()
        out = env.step(a, sample_next=True)
        ref_out = ref.step(_np(a), autoreset=True)
        assert_step_equal(env, out, ref_out, t)
        actions = env.next_action.clone()
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(env.stats), ref.stats()) and int(env.stats[0]) > 2 * n


@pytest.mark.parametrize("n,slots,segs,max_out", [(2048, 8, (128, 128, 60), None), (1000, 1, (90, 40), None), (4096 + 48, 3, (150,), None),
                                                   (48, 16, (300,), None), (2048, 8, (128, 100), "128"), (512, 8, (128,), "0"),
                                                   (2048, 8, (128, 60), "old-kernel")])
def test_mt19937_rollout_kernel_bit_exact(VecEnv, oracle, monkeypatch, n, slots, segs, max_out):
    """spl_rollout_random with the reference's own decks: shuffle='mt19937' + a ring of prefetched deals per env
    (spl_envs_t.spare_slots), refilled behind every launch.  Every output of every lock-step equals the oracle's (which
    deals with CPython's random.Random(seed).shuffle), whatever the ring size: with 8 slots a 128-step launch never runs
    out, with 1 or 3 slots envs do run out and are dealt in place by the kernel (the slow path).  The batch dealer
    (spl_spare_deal_kernel) leaves a deal to the consumer when it would need more than 227 generator outputs: lowering
    that limit to 128 (about half of all deals) or 0 (all) exercises the hand-over."""
    if max_out == "old-kernel":
        monkeypatch.setenv("SPL_DEAL_BATCH", "0")
    elif max_out is not None:
        monkeypatch.setenv("SPL_DEAL_MAX_OUTPUTS", max_out)
    env = VecEnv(n, seed=4242, shuffle="mt19937", autoreset=True, prefetch_deals=slots)
    assert env.spare_slots == slots
    ref = oracle.OracleVec(n, seed_base=4242)
    obs0, info0 = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(_np(obs0), robs) and np.array_equal(_np(info0["action_mask"]), rmask)
    actions = env.sample_random_actions().clone()
    t_abs = 0
    for T in segs:
        obs = torch.zeros((T, n, 297), dtype=torch.int32, device="cuda")
        mask = torch.zeros((T, n, 45), dtype=torch.int8, device="cuda")
        rew = torch.zeros((T, n), dtype=torch.float32, device="cuda")
        term = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        info = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        acts = torch.zeros((T + 1, n), dtype=torch.int32, device="cuda")
        acts[0] = actions
        env.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts, info=info)
        h_obs, h_mask, h_rew, h_term, h_info, h_acts = (_np(x) for x in (obs, mask, rew, term, info, acts))
        for t in range(T):
            robs, rrew, rterm, rinfo, rmask = ref.step(h_acts[t], autoreset=True)
            assert np.array_equal(h_obs[t], robs), f"step {t_abs}: obs"
            assert np.array_equal(h_mask[t], rmask), f"step {t_abs}: mask"
            assert np.array_equal(h_rew[t], rrew) and np.array_equal(h_term[t], rterm) and np.array_equal(h_info[t], rinfo), f"step {t_abs}"
            t_abs += 1
        actions = acts[T].clone()
        assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(env.stats), ref.stats()) and int(env.stats[0]) > n


@pytest.mark.parametrize("n,slots,T,segs", [(16384, 16, 64, 8), (16384, 2, 64, 8), (4096, 1, 40, 10)])
def test_mt19937_rollout_with_concurrent_refills(VecEnv, oracle, n, slots, T, segs):
    """rollout_random(refill=False) launches back to back on the main stream while refill_deals() runs on a side stream
    (scan + batch dealer concurrently with the NEXT launch).  The dealer publishes a row flag-last, the consumer reads the
    flag first; with 16 slots no env ever runs dry, with 1 or 2 the rings do run dry while they are being refilled -- envs
    are then dealt in place, slots end up with stale tags and are re-dealt by a later scan.  Whatever the interleaving,
    every output equals the oracle's."""
    env = VecEnv(n, seed=777, shuffle="mt19937", autoreset=True, prefetch_deals=slots)
    ref = oracle.OracleVec(n, seed_base=777)
    env.reset()
    ref.reset()
    TT = T * segs
    obs = torch.zeros((TT, n, 297), dtype=torch.int32, device="cuda")
    mask = torch.zeros((TT, n, 45), dtype=torch.int8, device="cuda")
    rew = torch.zeros((TT, n), dtype=torch.float32, device="cuda")
    term = torch.zeros((TT, n), dtype=torch.uint8, device="cuda")
    info = torch.zeros((TT, n), dtype=torch.uint8, device="cuda")
    acts = torch.zeros((TT + 1, n), dtype=torch.int32, device="cuda")
    acts[0] = env.sample_random_actions()
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    pending = []
    for k in range(segs):  # no host synchronisation inside this loop
        if len(pending) >= 2:
            main.wait_event(pending[-2])
        lo, hi = k * T, (k + 1) * T
        env.rollout_random(T, acts[lo], obs=obs[lo:hi], mask=mask[lo:hi], reward=rew[lo:hi], terminated=term[lo:hi],
                           next_actions=acts[lo:hi + 1], info=info[lo:hi], refill=False)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            env.refill_deals()
            ev = torch.cuda.Event()
            ev.record(side)
        pending.append(ev)
    main.wait_stream(side)
    torch.cuda.synchronize()
    h_obs, h_mask, h_rew, h_term, h_info, h_acts = (_np(x) for x in (obs, mask, rew, term, info, acts))
    for t in range(TT):
        robs, rrew, rterm, rinfo, rmask = ref.step(h_acts[t], autoreset=True)
        assert np.array_equal(h_obs[t], robs), f"step {t}: obs"
        assert np.array_equal(h_mask[t], rmask), f"step {t}: mask"
        assert np.array_equal(h_rew[t], rrew) and np.array_equal(h_term[t], rterm) and np.array_equal(h_info[t], rinfo), f"step {t}"
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(env.stats), ref.stats()) and int(env.stats[0]) > n
    env.refill_deals()
    torch.cuda.synchronize()
    rows = env.spare[: n * slots * 96].view(n * slots, 96)
    assert bool((rows[:, 95] == 1).all())  # every slot ready again


def test_mt19937_rollout_then_steps_share_the_ring(VecEnv):
    """A ring of prefetched deals serves spl_step and spl_rollout_random alike: interleaving them gives the trajectory of
    plain chained steps without any prefetching (the in-line reset kernel path pinned against the oracle above)."""
    n = 1024 + 32
    a = VecEnv(n, seed=6, shuffle="mt19937", autoreset=True, prefetch_deals=4)
    b = VecEnv(n, seed=6, shuffle="mt19937", autoreset=True, prefetch_deals=False)
    a.reset()
    b.reset()
    actions = a.sample_random_actions().clone()
    for rep in range(3):
        T = 50
        obs = torch.zeros((T, n, 297), dtype=torch.int32, device="cuda")
        mask = torch.zeros((T, n, 45), dtype=torch.int8, device="cuda")
        rew = torch.zeros((T, n), dtype=torch.float32, device="cuda")
        term = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        acts = torch.zeros((T + 1, n), dtype=torch.int32, device="cuda")
        acts[0] = actions
        a.rollout_random(T, acts[0], obs=obs, mask=mask, reward=rew, terminated=term, next_actions=acts)
        for t in range(T):
            ob, rb, tb, _, _ = b.step(acts[t].clone(), sample_next=True)
            assert torch.equal(ob, obs[t]) and torch.equal(b.mask, mask[t]) and torch.equal(rb, rew[t]), f"rep {rep} step {t}"
            assert torch.equal(b.next_action, acts[t + 1])
        actions = acts[T].clone()
        for t in range(40):
            oa, ra, ta, _, _ = a.step(actions, sample_next=True)
            ob, rb, tb, _, _ = b.step(actions, sample_next=True)
            assert torch.equal(oa, ob) and torch.equal(a.mask, b.mask) and torch.equal(ra, rb) and torch.equal(ta, tb), f"rep {rep} step {t}"
            assert torch.equal(a.next_action, b.next_action)
            actions = a.next_action.clone()
        if rep == 1:
            m = torch.zeros(n, dtype=torch.bool, device="cuda")
            m[::3] = True
            a.reset(reset_mask=m)
            b.reset(reset_mask=m)
            assert torch.equal(a.obs, b.obs)
            actions = a.sample_random_actions().clone()
    assert torch.equal(a.export_state(), b.export_state()) and torch.equal(a.stats, b.stats) and torch.equal(a.episode, b.episode)


def test_mt19937_manual_refills_with_a_frozen_step_counter(VecEnv, oracle, monkeypatch):
    """A caller whose io->action_t never reaches the library's refill cadence (a replayed single-step CUDA graph) refills
    the rings itself with spl_refill_spares; outputs stay those of the oracle and (almost) no env is dealt in place."""
    monkeypatch.setenv("SPL_SPARE_REFILL_AGE", "1000000")  # the library's own cadence never fires
    n, steps = 2048, 200
    env = VecEnv(n, seed=31, shuffle="mt19937", autoreset=True, prefetch_deals=2)
    ref = oracle.OracleVec(n, seed_base=31)
    env.reset()
    ref.reset()
    actions = env.sample_random_actions().clone()
    for t in range(steps):
        a = actions.clone()
        out = env.step(a, sample_next=True)
        ref_out = ref.step(_np(a), autoreset=True)
        assert_step_equal(env, out, ref_out, t)
        actions = env.next_action.clone()
        if t % 30 == 29:
            env.refill_deals()
            torch.cuda.synchronize()
            # every slot is ready again: ready byte (95) of each 96-byte row
            rows = env.spare[: n * 2 * 96].view(n * 2, 96)
            assert bool((rows[:, 95] == 1).all())
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert np.array_equal(_np(env.stats), ref.stats()) and int(env.stats[0]) > n


def test_mt19937_prefetched_deals_survive_manual_resets(VecEnv):
    """Masked manual resets re-deal the chosen envs AND their spares: an env with prefetched deals stays identical to one
    without (the in-line reset path, which the oracle tests pin) through auto-resets, masked resets and a full reset."""
    n = 1024 + 32
    a = VecEnv(n, seed=5, shuffle="mt19937", autoreset=True, prefetch_deals=True)
    b = VecEnv(n, seed=5, shuffle="mt19937", autoreset=True, prefetch_deals=False)
    a.reset()
    b.reset()
    actions = a.sample_random_actions().clone()
    for t in range(300):
        oa, ra, ta, _, _ = a.step(actions, sample_next=True)
        ob, rb, tb, _, _ = b.step(actions, sample_next=True)
        assert torch.equal(oa, ob) and torch.equal(a.mask, b.mask) and torch.equal(ra, rb) and torch.equal(ta, tb), f"step {t}"
        assert torch.equal(a.info_bits, b.info_bits) and torch.equal(a.next_action, b.next_action)
        actions = a.next_action.clone()
        if t in (70, 71, 150):
            m = torch.zeros(n, dtype=torch.bool, device="cuda")
            m[t % 5::5] = True
            a.reset(reset_mask=m)
            b.reset(reset_mask=m)
            assert torch.equal(a.obs, b.obs) and torch.equal(a.mask, b.mask)
            actions = a.sample_random_actions().clone()
        if t == 220:
            a.reset()
            b.reset()
            actions = a.sample_random_actions().clone()
    assert torch.equal(a.export_state(), b.export_state()) and torch.equal(a.stats, b.stats) and torch.equal(a.episode, b.episode)
    assert int(a.stats[0]) > 2 * n


def test_fused_reset_equals_reset_kernel(VecEnv):
    """The in-step Philox deal and spl_reset(reset_mask=...) produce the same new episode for (seed, env, episode)."""
    n, T = 4096, 120
    a = VecEnv(n, seed=8, shuffle="philox", autoreset=True)
    b = VecEnv(n, seed=8, shuffle="philox", autoreset=False)
    a.reset()
    b.reset()
    actions = a.sample_random_actions().clone()
    resets = 0
    for t in range(T):
        a.step(actions, sample_next=True)
        b.step(actions)
        done = b.terminated.clone()
        if bool(done.any()):
            b.reset(reset_mask=done)
            resets += int(done.sum())
        assert torch.equal(a.obs, b.obs) and torch.equal(a.mask, b.mask), f"step {t}"
        actions = a.next_action.clone()
    assert resets > 1000
    assert torch.equal(a.export_state(), b.export_state()) and torch.equal(a.episode, b.episode)


def test_unaligned_output_buffers_and_strided_state(VecEnv, oracle):
    """The C ABI accepts caller buffers that are not 16-byte aligned (scalar store path) and a state stride larger
    than n (a shard that is a prefix of a bigger allocation); results are unchanged."""
    import ctypes as C

    from splendor_gym_b200 import _lib as L

    n, cap = 200, 256
    env = VecEnv(cap, seed=3, shuffle="mt19937", autoreset=True)
    # shrink the logical shard to the first n envs of the cap-sized planes
    env._envs.n = n
    env.n = n
    ref = oracle.OracleVec(n, seed_base=3)
    big_obs = torch.zeros(cap * 297 + 8, dtype=torch.int32, device="cuda")
    big_mask = torch.zeros(cap * 45 + 8, dtype=torch.int8, device="cuda")
    obs = big_obs[1:1 + n * 297].view(n, 297)      # 4-byte offset: not 16-byte aligned
    mask = big_mask[3:3 + n * 45].view(n, 45)      # 3-byte offset
    assert obs.data_ptr() % 16 != 0 and mask.data_ptr() % 16 != 0
    lib = L.load()
    stream = torch.cuda.current_stream().cuda_stream
    L.check(lib.spl_reset(C.byref(env._envs), None, None, obs.data_ptr(), mask.data_ptr(), stream))
    robs, rmask = ref.reset()
    assert np.array_equal(_np(obs), robs) and np.array_equal(_np(mask), rmask)
    env._is_reset = True
    for t in range(150):
        a = ref.random_actions(7, t)
        out = env.step(torch.from_numpy(a).cuda(), out_obs=obs, out_mask=mask)
        robs, rrew, rterm, rinfo, rmask = ref.step(a, autoreset=True)
        assert np.array_equal(_np(obs), robs), f"step {t}"
        assert np.array_equal(_np(mask), rmask), f"step {t}"
        assert np.array_equal(_np(env.reward)[:n], rrew) and np.array_equal(_np(env._terminated)[:n], rterm)
    assert int(big_obs[0]) == 0 and int(big_obs[1 + n * 297:].abs().sum()) == 0  # nothing written outside the view
    assert int(big_mask[:3].abs().sum()) == 0 and int(big_mask[3 + n * 45:].abs().sum()) == 0
    rows = _np(env.export_state())[:n]
    assert np.array_equal(rows, ref.export_rows())


def _gpu_digest_games(VecEnv, seeds, T):
    """Play the LCG-policy games of `seeds` on the GPU, one env per game in lock-step (policy on the device from the kernel's
    own masks); returns (moves, winner, steps, per-game sha256) -- see tests/digest_util.py for the record layout."""
    import digest_util as D

    n = len(seeds)
    env = VecEnv(n, shuffle="mt19937", autoreset=False)
    _, info = env.reset(seeds=torch.from_numpy(seeds.astype(np.int64)))
    dev = env.device
    x = torch.from_numpy(D.lcg_seed(seeds).astype(np.int64)).to(dev)
    mask = info["action_mask"]
    active = torch.ones(n, dtype=torch.bool, device=dev)
    steps = torch.zeros(n, dtype=torch.int64, device=dev)
    recs = torch.zeros((T, n, D.REC), dtype=torch.uint8, device=dev)
    for t in range(T):
        x = (1664525 * x + 1013904223) % (1 << 32)
        cnt = mask.sum(1, dtype=torch.int64)
        k = (x >> 16) % cnt.clamp(min=1)
        a = (mask.to(torch.int64).cumsum(1) > k[:, None]).to(torch.uint8).argmax(1).to(torch.int32)
        a[cnt == 0] = 0
        obs, rew, term, trunc, info = env.step(a, active=active)
        mask = info["action_mask"]
        r = recs[t]
        r[:, :297] = obs.to(torch.uint8)
        r[:, 297:342] = mask.view(torch.uint8)
        r[:, 342:346] = rew.view(torch.uint8).view(n, 4)
        r[:, 346] = term.to(torch.uint8)
        r[:, 347] = info["info_bits"]
        steps += active
        active = active & ~term.bool()
    assert not bool(active.any())
    assert int(env.obs.max()) < 256
    rows = _np(env.export_state())
    st = _np(steps)
    return rows[:, 72].copy(), rows[:, 74].copy(), st, D.game_shas(_np(recs), st)


def test_ten_thousand_reference_games_by_digest(VecEnv):
    """CUDA directly against the reference at volume: the 10,000 games of tests/golden/games_digest.json (773,764 env-steps
    played by the unmodified Python engine under the LCG policy) replayed on the GPU.  Every step's observation, mask, reward,
    terminated flag and info bits go into the game's sha256, which must equal the reference's, as must move counts and winners."""
    G = load_golden("games_digest.json")
    games = np.array([g[:3] for g in G["games"]], np.int64)
    n = len(games)
    moves, winner, st, shas = _gpu_digest_games(VecEnv, G["seed0"] + np.arange(n, dtype=np.int64), int(games[:, 2].max()))
    assert np.array_equal(st, games[:, 2]) and np.array_equal(moves, games[:, 0]) and np.array_equal(winner, games[:, 1])
    bad = [i for i in range(n) if shas[i] != G["games"][i][3]]
    assert not bad, f"{len(bad)} of {n} game digests differ from the reference, first: seed {G['seed0'] + bad[0]}"


def test_hundred_thousand_reference_games_by_chunk_digest(VecEnv):
    """... and the 100,000 games (7,732,816 env-steps) of tests/golden/games_digest_100k.json, in slices of 25,000 envs."""
    import digest_util as D

    G = load_golden("games_digest_100k.json")
    got, total = [], 0
    for lo in range(0, G["games"], 25000):
        seeds = G["seed0"] + np.arange(lo, min(G["games"], lo + 25000), dtype=np.int64)
        moves, winner, st, shas = _gpu_digest_games(VecEnv, seeds, G["max_steps"])
        got += D.chunk_digests(moves, winner, st, shas, G["chunk"])
        total += int(st.sum())
    bad = [i for i, (a, b) in enumerate(zip(got, G["chunks"])) if a != b]
    assert not bad and len(got) == len(G["chunks"]), f"{len(bad)} of {len(got)} chunk digests differ from the reference, first chunk {bad[:1]}"
    assert total == G["env_steps"]


def test_rollout_without_observations_matches_oracle(VecEnv, oracle):
    """BASELINE configs[3] (mask + step only): write_obs=False on the single-step path and obs=None on the rollout
    kernel produce the same masks, rewards, terminations, info bits, actions and states as the oracle."""
    n, T = 4096 + 17, 48
    env = VecEnv(n, seed=77, shuffle="mt19937", autoreset=True, prefetch_deals=8)
    ref = oracle.OracleVec(n, seed_base=77)
    env.reset()
    ref.reset()
    dev = env.device
    actions = env.sample_random_actions().clone()
    for t in range(20):  # one launch per lock-step, no observation
        a = _np(actions).copy()
        obs, rew, term, trunc, info = env.step(actions, sample_next=True, write_obs=False)
        robs, rrew, rterm, rinfo, rmask = ref.step(a, autoreset=True)
        assert np.array_equal(_np(info["action_mask"]), rmask), f"step {t}: mask"
        assert np.array_equal(_np(rew), rrew) and np.array_equal(_np(term).astype(np.uint8), rterm) and np.array_equal(_np(info["info_bits"]), rinfo)
        actions = env.next_action.clone()
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    mask = torch.zeros((T, n, 45), dtype=torch.int8, device=dev)
    rew = torch.zeros((T, n), dtype=torch.float32, device=dev)
    term = torch.zeros((T, n), dtype=torch.uint8, device=dev)
    info_b = torch.zeros((T, n), dtype=torch.uint8, device=dev)
    nxt = torch.zeros((T + 1, n), dtype=torch.int32, device=dev)
    env.rollout_random(T, actions, obs=None, mask=mask, reward=rew, terminated=term, next_actions=nxt, info=info_b)
    a = _np(actions)
    for t in range(T):
        robs, rrew, rterm, rinfo, rmask = ref.step(a, autoreset=True)
        assert np.array_equal(_np(mask[t]), rmask), f"rollout step {t}: mask"
        assert np.array_equal(_np(rew[t]), rrew) and np.array_equal(_np(term[t]), rterm) and np.array_equal(_np(info_b[t]), rinfo)
        a = _np(nxt[t + 1])
    assert np.array_equal(_np(env.export_state()), ref.export_rows())
    assert int(env.stats[0]) > 0


@pytest.mark.parametrize("how", ["seed_table", "seed_table_ring1", "seed_table_no_ring", "load_deals"])
def test_replay_of_reference_autoreset_stream(VecEnv, how):
    """Replay mode (north_star: "accepts the reference's deck permutations").  tests/golden/autoreset_stream.json holds
    reference SplendorEnv instances that were reset(seed=s) and then auto-reset four times, each time drawing a new engine
    seed from their own PCG64 stream.  Fed with those draws (set_episode_seeds) or with the dealt decks themselves
    (load_deals), the batched path reproduces every env step by step across all the auto-resets."""
    from test_oracle_golden import digest

    G = load_golden("autoreset_stream.json")
    n = len(G)
    prefetch = {"seed_table": 4, "seed_table_ring1": 1, "seed_table_no_ring": False, "load_deals": 4}[how]
    env = VecEnv(n, shuffle="mt19937", autoreset=True, prefetch_deals=prefetch, seed=555)
    if how != "load_deals":
        env.set_episode_seeds(torch.tensor([g["engine_seeds"][1:] for g in G], dtype=torch.int64))
    env.reset(seeds=torch.tensor([g["engine_seeds"][0] for g in G], dtype=torch.int64))
    if how == "load_deals":
        starts = torch.tensor([g["starts"][1:] for g in G], dtype=torch.int32)  # [n, 4, 166]
        env.load_deals(torch.stack([VecEnv.deals_from_rows(starts[:, k]) for k in range(4)], dim=1))
        with pytest.raises(Exception, match="outside the engine's domain"):
            bad = torch.zeros((n, 1, 96), dtype=torch.uint8)
            env.load_deals(bad)
    assert np.array_equal(_np(env.export_state()), np.array([g["starts"][0] for g in G], np.int32))
    T = max(len(g["actions"]) for g in G)
    for t in range(T):
        live = np.array([t < len(g["actions"]) for g in G])
        a = np.array([g["actions"][t] if live[i] else 0 for i, g in enumerate(G)], np.int32)
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda(), active=torch.from_numpy(live).cuda())
        rows = _np(env.export_state())
        o, m, r, te, ib = _np(obs), _np(info["action_mask"]), _np(rew), _np(term), _np(info["info_bits"])
        for i, g in enumerate(G):
            if live[i]:
                assert digest(o[i], m[i], rows[i], r[i], te[i], int(ib[i])) == g["digests"][t], (how, g["seed"], t)
    assert _np(env.episode).tolist() == [4] * n


def test_paired_warp_step_kernel_is_bit_exact(VecEnv, oracle, monkeypatch):
    """The experimental two-warps-per-tile single-step kernel (SPL_STEP_PAIRED=1: rules warpgroup + helper warpgroup with
    setmaxnreg register hand-over) against the oracle: every output of every step, Philox and MT19937 decks, ragged size."""
    monkeypatch.setenv("SPL_STEP_PAIRED", "1")
    for shuffle, n in (("mt19937", 4096 + 33), ("philox", 1000)):
        env = VecEnv(n, seed=31, shuffle=shuffle, autoreset=True)  # (spl_init re-reads the knobs)
        obs, info = env.reset()
        ref = oracle.OracleVec(n, seed_base=31)
        if shuffle == "mt19937":
            ref.reset()
        else:
            ref.import_rows(_np(env.export_state()))
        actions = env.sample_random_actions().clone()
        for t in range(90):
            a = _np(actions).copy()
            out = env.step(actions, sample_next=True)  # (Philox decks: the oracle is handed the GPU's state of every re-dealt game)
            if shuffle == "mt19937":
                assert_step_equal(env, out, tuple(x.copy() for x in ref.step(a, autoreset=True)), t, check_state=ref.export_rows())
            else:
                robs, rrew, rterm, rinfo, rmask = (x.copy() for x in ref.step(a, autoreset=False))
                assert np.array_equal(_np(out[1]), rrew) and np.array_equal(_np(out[2]).astype(np.uint8), rterm), t
                rows = _np(env.export_state())
                done = rterm.astype(bool)
                assert np.array_equal(rows[~done], ref.export_rows()[~done]), t
                assert np.array_equal(_np(out[0])[~done], robs[~done]) and np.array_equal(_np(out[4]["action_mask"])[~done], rmask[~done]), t
                ref.import_rows(rows, which=done.astype(np.uint8))
            actions = env.next_action.clone()
    monkeypatch.setenv("SPL_STEP_PAIRED", "0")
    VecEnv(64, seed=1)  # restore the default knobs for the tests that follow
