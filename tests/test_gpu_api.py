"""The reference's single-environment API surface (SplendorEnv, engine functions, wrappers) on top of the CUDA
kernels.  These tests restate what the reference's own test-suite pins (splendor_gym/tests/*.py; SURVEY.md
section 4), against this package's import paths, and tie the facade to the batched path."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def first_legal(mask):
    legal = np.flatnonzero(mask)
    return int(legal[0]) if len(legal) else 0


# ----------------------------------------------------------------------------- engine functions
def test_initial_state_matches_reference_fingerprints():
    """initial_state(seed) boards / nobles / deck tops / obs hash as produced by the reference."""
    from splendor_gym_b200.engine import initial_state, legal_moves
    from splendor_gym_b200.engine.encode import TOTAL_ACTIONS, encode_observation

    want = {
        0: ([[24, 26, 2, 16], [54, 67, 56, 48], [86, 85, 73, 79]], [1007, 1009, 1005], [32, 41, 89], "acccb83d866e5f0c"),
        42: ([[7, 1, 17, 15], [51, 67, 69, 59], [81, 75, 89, 87]], [1004, 1009, 1001], [14, 48, 76], "1e1673d87f5d6bc3"),
    }
    for seed, (board, nobles, tops, sha) in want.items():
        s = initial_state(seed=seed)
        assert [[c.id for c in s.board[t]] for t in (1, 2, 3)] == board
        assert [n.id for n in s.nobles] == nobles
        assert [s.decks[t][-1].id for t in (1, 2, 3)] == tops
        assert hashlib.sha256(encode_observation(s).tobytes()).hexdigest()[:16] == sha
        m = legal_moves(s)
        assert len(m) == TOTAL_ACTIONS and any(m)
        assert s.bank == [4, 4, 4, 4, 4, 5] and s.to_play == 0 and s.turn_count == 1 and s.move_count == 0


def test_take_actions_move_tokens_and_keep_turn_count():
    from splendor_gym_b200.engine import apply_action, initial_state, legal_moves

    s = initial_state(seed=1)
    a = first_legal(legal_moves(s))
    before = sum(s.bank)
    s2 = apply_action(s, a)
    assert a < 10 and sum(s2.bank) == before - 3
    assert s2.turn_count == s.turn_count and s2.move_count == 1 and s2.to_play == 1
    assert sum(s.bank) == before  # apply_action is pure: the input state is untouched
    s3 = apply_action(s, 12)
    assert sum(s3.bank) == before - 2 and s3.players[0].tokens[2] == 2
    with pytest.raises(ValueError):
        apply_action(s, 45)


def test_mask_invariants():
    from splendor_gym_b200.engine import initial_state, legal_moves

    s = initial_state(seed=7)
    s.bank[1] = 3
    s.board[2][3] = None
    s.decks[3].clear()
    m = legal_moves(s)
    p = s.players[s.to_play]
    for ci in range(5):
        assert (m[10 + ci] == 1) == (s.bank[ci] >= 4)
    for tier in (1, 2, 3):
        assert (m[39 + tier - 1] == 1) == (len(s.decks[tier]) > 0 and len(p.reserved) < 3)
        for slot in range(4):
            if s.board[tier][slot] is None:
                assert m[15 + (tier - 1) * 4 + slot] == 0 and m[27 + (tier - 1) * 4 + slot] == 0
    assert m[15 + 4 + 3] == 0 and m[27 + 4 + 3] == 0 and m[41] == 0


def test_token_limit_after_oversized_hand():
    """Hands above the limit (as the reference's tests inject them) come back to exactly 10, non-gold first."""
    from splendor_gym_b200.engine import apply_action, initial_state

    s = initial_state(seed=0)
    s.players[0].tokens = [5, 5, 5, 5, 5, 0]
    s2 = apply_action(s, 0)
    assert sum(s2.players[0].tokens) == 10
    g = next(c for c in load_golden("edge_cases.json") if c["name"] == "token_limit_25_tokens")
    assert s2.players[0].tokens == g["row_out"][6:12] and s2.bank == g["row_out"][0:6]


# ----------------------------------------------------------------------------- SplendorEnv facade
def test_env_reset_step_types_and_seeding():
    from splendor_gym_b200.envs import SplendorEnv

    env = SplendorEnv(num_players=2)
    assert env.action_space.n == 45 and env.observation_space.shape == (297,)
    for rec in load_golden("env_seeding.json"):  # gymnasium's PCG64 seeding path, pinned from the reference run
        obs, info = env.reset(seed=rec["seed"])
        assert obs.shape == (297,) and obs.dtype == np.int32
        assert info["action_mask"].shape == (45,) and info["action_mask"].dtype == np.int8 and info["to_play"] == 0
        assert [c.id for c in env.state.board[1]] == rec["tier1_board"]
    obs, r, term, trunc, info = env.step(first_legal(info["action_mask"]))
    assert isinstance(r, float) and isinstance(term, bool) and trunc is False and "action_mask" in info
    with pytest.raises(NotImplementedError):
        SplendorEnv(num_players=3)


def test_env_determinism_and_random_play():
    from splendor_gym_b200.envs import SplendorEnv

    e1, e2 = SplendorEnv(), SplendorEnv()
    o1, i1 = e1.reset(seed=42)
    o2, i2 = e2.reset(seed=42)
    assert np.array_equal(o1, o2)
    rng = np.random.RandomState(0)
    for _ in range(50):
        legal = np.flatnonzero(i1["action_mask"])
        a = int(legal[rng.randint(len(legal))]) if len(legal) else 0
        o1, r1, t1, _, i1 = e1.step(a)
        o2, r2, t2, _, i2 = e2.step(a)
        assert np.array_equal(o1, o2) and (r1, t1) == (r2, t2) and np.array_equal(i1["action_mask"], i2["action_mask"])
        if t1:
            break


def test_env_illegal_action_and_errors():
    from splendor_gym_b200.envs import SplendorEnv

    env = SplendorEnv()
    with pytest.raises(AssertionError):
        env.step(0)
    obs, info = env.reset(seed=3)
    illegal = int(np.flatnonzero(info["action_mask"] == 0)[0])
    before = env.state.move_count
    obs2, r, term, trunc, info2 = env.step(illegal)
    assert r == pytest.approx(-0.01) and not term and info2.get("illegal_action") and env.state.move_count == before
    assert np.array_equal(obs, obs2) and np.array_equal(info["action_mask"], info2["action_mask"])
    with pytest.raises(ValueError):
        env.step(45)
    env.state.game_over = True
    with pytest.raises(RuntimeError):
        env.step(0)


@pytest.mark.gpu
def test_vec_env_check_raises_like_the_reference():
    """Batched error convention: where SplendorEnv.step raises, the lock-step records SPL_INFO_ERROR and leaves
    the env untouched; check=True turns that into the reference's exceptions."""
    import torch
    from splendor_gym_b200 import SplendorVecEnv

    env = SplendorVecEnv(64, seed=3, shuffle="mt19937", autoreset=False)
    env.reset()
    a = env.sample_random_actions().clone()
    env.step(a, check=True)  # legal actions: nothing raised
    before = env.export_state().clone()
    a = env.sample_random_actions().clone()
    a[7] = 45
    a[9] = -1
    with pytest.raises(ValueError, match="env 7: action 45"):
        env.step(a, check=True)
    after = env.export_state()
    assert torch.equal(before[7], after[7]) and torch.equal(before[9], after[9])  # untouched, as after the reference's raise
    # play env 0 to the end without auto-reset, then step it again
    for _ in range(400):
        env.step(env.sample_random_actions())
        if bool(env.terminated.all()):
            break
    assert bool(env.terminated.all())
    with pytest.raises(RuntimeError, match="terminal"):
        env.step(torch.zeros(64, dtype=torch.int32, device="cuda"), check=True)


def test_env_no_legal_move_draw():
    from splendor_gym_b200.engine import legal_moves
    from splendor_gym_b200.envs import SplendorEnv

    env = SplendorEnv()
    env.reset(seed=0)
    env.state.bank[:] = [0, 0, 0, 0, 0, 0]
    p = env.state.players[env.state.to_play]
    p.tokens[:] = [10, 0, 0, 0, 0, 0]
    p.reserved = env.state.decks[1][:3]
    for t in (1, 2, 3):
        env.state.board[t] = [None, None, None, None]
    assert not any(legal_moves(env.state))
    obs, r, term, trunc, info = env.step(0)
    assert term and r == 0 and env.state.winner_index is None and info.get("draw") and "final_rewards" not in info
    assert not info["action_mask"].any()


def test_env_reduced_take3():
    from splendor_gym_b200.engine import legal_moves
    from splendor_gym_b200.envs import SplendorEnv

    env = SplendorEnv()
    env.reset(seed=123)
    env.state.bank[:] = [1, 0, 2, 0, 0, 0]
    m = legal_moves(env.state)
    assert [i for i in range(10) if m[i]] == [0, 3, 4]
    env.step(0)
    last = env.state.players[1 - env.state.to_play]
    assert last.tokens[0] + last.tokens[2] == 2
    env.reset(seed=123)
    env.state.bank[:] = [0, 0, 0, 0, 3, 0]
    m = legal_moves(env.state)
    assert [i for i in range(10) if m[i]] == [2, 4, 5, 7, 8, 9]
    env.step(2)
    assert env.state.players[1 - env.state.to_play].tokens[4] == 1


def test_observation_layout_and_hidden_reservation():
    """Section offsets (bank 0:6, me 6:19, opponent 19:32, board 32:188, reserved 188:272, nobles 272:290, tail
    290:297); my reserved cards always carry revealed=1; the opponent's blind reservation is 14 zeros."""
    from splendor_gym_b200.envs import SplendorEnv

    env = SplendorEnv()
    obs, info = env.reset(seed=9)
    assert obs[0:6].tolist() == [4, 4, 4, 4, 4, 5] and obs[6:32].sum() == 0
    assert obs[290:293].tolist() == [36, 26, 16] and obs[293:297].tolist() == [1, 0, 0, 0]
    assert all(obs[32 + 13 * k] == 1 for k in range(12)) and all(obs[272 + 6 * k] == 1 for k in range(3))
    obs, *_ = env.step(39)  # P0 reserves blind (tier 1)
    assert obs[18 + 13] == 1 and obs[230:244].sum() == 0  # P1 sees the count, not the card
    obs, *_ = env.step(27)  # P1 reserves the visible tier-1 slot 0
    own = obs[188:202]
    assert own[0] == 1 and own[13] == 1 and own[1] == 1  # P0 sees its own hidden card, tier 1
    theirs = obs[230:244]
    assert theirs[0] == 1 and theirs[13] == 1  # P1's reservation came from the board: public


# ----------------------------------------------------------------------------- wrappers
def test_selfplay_and_dual_step_wrappers_against_batched_path():
    """Single-env wrappers driven by Python callbacks == SplendorVecEnv.dual_step with the same policies."""
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200.envs import SplendorEnv
    from splendor_gym_b200.wrappers import DualStepNativeWrapper, SelfPlayWrapper, vec_selfplay_step

    def opp(obs, info):  # deterministic opponent: highest legal action index
        legal = np.flatnonzero(info["action_mask"])
        return int(legal[-1]) if len(legal) else 0

    def vec_opp(obs, mask):
        idx = torch.arange(45, device=mask.device, dtype=torch.int32)
        return torch.where(mask > 0, idx, torch.full_like(idx, -1)).max(dim=1).values.clamp(min=0).to(torch.int32)

    seeds = [11, 12, 13, 14]
    engine_seeds = [int(np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))).integers(0, 2**31 - 1)) for s in seeds]
    for mode in ("native", "selfplay"):
        vec = SplendorVecEnv(len(seeds), shuffle="mt19937", autoreset=False)
        vobs, vinfo = vec.reset(seeds=torch.tensor(engine_seeds, dtype=torch.int64))
        singles = []
        for s in seeds:
            base = SplendorEnv()
            w = (DualStepNativeWrapper if mode == "native" else SelfPlayWrapper)(base, opponent_policy=opp, random_starts=True)
            o, i = w.reset(seed=s)
            singles.append([w, o, i, False])
        for it in range(80):
            acts = [first_legal(x[2]["action_mask"]) for x in singles]
            alive = [not x[3] for x in singles]
            if not any(alive):
                break
            a = torch.tensor(acts, dtype=torch.int32, device="cuda")
            if mode == "native":
                o, ar, _, orr, done, _ = vec.dual_step(a, vec_opp)
            else:
                o, ar, done, _, _ = vec_selfplay_step(vec, a, vec_opp)
            for k, x in enumerate(singles):
                if x[3]:
                    continue
                if mode == "native":
                    so, sr, _, sor, sd, si = x[0].dual_step(acts[k])
                    assert float(orr[k]) == pytest.approx(sor)
                else:
                    so, sr, sd, _, si = x[0].step(acts[k])
                assert np.array_equal(o[k].cpu().numpy(), so) and float(ar[k]) == pytest.approx(sr) and bool(done[k]) == sd
                assert np.array_equal(vec.mask[k].cpu().numpy(), si["action_mask"])
                x[1], x[2], x[3] = so, si, sd
        assert all(x[3] for x in singles)


def _det_pick(obs, mask, mul, add):
    """oracle/gen_golden.py:pick -- k-th legal action with k = (sum(obs) * mul + add) mod #legal (what the fixture's players do)."""
    legal = np.flatnonzero(np.asarray(mask))
    return 0 if len(legal) == 0 else int(legal[(int(np.asarray(obs).sum()) * mul + add) % len(legal)])


def _sha12(a, dtype):
    return hashlib.sha256(np.ascontiguousarray(a, dtype).tobytes()).hexdigest()[:12]


def test_wrappers_match_reference_generated_turns():
    """tests/golden/wrappers.json was produced by the reference's SelfPlayWrapper / DualStepNativeWrapper /
    DualStepSelfPlayWrapper (51 games incl. turn-limit draws -- +0.1 under SelfPlayWrapper, -0.1 under the dual-step wrappers
    -- and a no-legal-move draw on the agent's move).  The three single-env wrappers here (adapters over the device turn) must
    return the same observation, mask, reward, done flag, opponent move / reward and bookkeeping at every step."""
    from splendor_gym_b200.engine.state import row_to_state, state_to_row
    from splendor_gym_b200.envs import SplendorEnv
    from splendor_gym_b200.wrappers import DualStepNativeWrapper, DualStepSelfPlayWrapper, SelfPlayWrapper

    kinds = {"selfplay": SelfPlayWrapper, "dual_native": DualStepNativeWrapper, "dual_selfplay": DualStepSelfPlayWrapper}
    seen_rewards = set()
    for g in load_golden("wrappers.json"):
        env = kinds[g["wrapper"]](SplendorEnv(), opponent_policy=lambda o, i: _det_pick(o, i["action_mask"], 7, 3), random_starts=False)
        env.reset(seed=g["seed"])
        env.env.state = row_to_state(np.array(g["row0"], np.int32))
        for t, want in enumerate(g["steps"]):
            obs, r, done, trunc, info = env.step(want["a"])
            where = (g["wrapper"], g["seed"], t)
            assert r == pytest.approx(want["r"]) and done == want["done"] and trunc is False, where
            assert _sha12(obs, np.int32) == want["obs"] and _sha12(info["action_mask"], np.int8) == want["mask"], where
            assert _sha12(state_to_row(env.env.state), np.int32) == want["row"], where
            for k in ("opponent_action", "game_ended_on", "phase", "turn_limit", "draw"):
                if k in want:
                    assert info.get(k) == want[k], (where, k)
                elif k in ("turn_limit", "draw"):
                    assert k not in info, (where, k)
            if "opponent_reward" in want:
                assert info["opponent_reward"] == pytest.approx(want["opponent_reward"]), where
            if "final_rewards" in want:
                assert [info["final_rewards"][0], info["final_rewards"][1]] == pytest.approx(want["final_rewards"]), where
            if done:
                seen_rewards.add((g["wrapper"], round(float(r), 2), bool(want.get("turn_limit"))))
        assert done
        if "stats" in g:
            st = env.get_wrapper_stats()
            assert {k: st[k] for k in g["stats"]} == g["stats"]
    assert ("selfplay", 0.1, True) in seen_rewards and ("dual_selfplay", -0.1, True) in seen_rewards and ("dual_native", -0.1, True) in seen_rewards


def test_batched_dual_step_matches_reference_generated_turns():
    """The same fixture through the BATCHED path: all 51 games side by side in one SplendorVecEnv, a turn =
    dual_step(reward_mode=...) with the opponent's deterministic policy evaluated on the device."""
    from splendor_gym_b200 import SplendorVecEnv

    for mode, kinds in (("selfplay", ("selfplay",)), ("native", ("dual_native", "dual_selfplay"))):
        games = [g for g in load_golden("wrappers.json") if g["wrapper"] in kinds]
        n = len(games)
        vec = SplendorVecEnv(n, shuffle="mt19937", autoreset=False)
        vec.reset()
        vec.import_state(torch.tensor([g["row0"] for g in games], dtype=torch.int32))
        moves = {}

        def opponent(obs, mask):
            cnt = mask.sum(1, dtype=torch.int64)
            k = (obs.sum(1, dtype=torch.int64) * 7 + 3) % cnt.clamp(min=1)
            a = (mask.to(torch.int64).cumsum(1) > k[:, None]).to(torch.uint8).argmax(1).to(torch.int32)
            a[cnt == 0] = 0
            moves["opp"] = a
            return a

        for t in range(max(len(g["steps"]) for g in games)):
            live = [t < len(g["steps"]) for g in games]
            acts = torch.tensor([g["steps"][t]["a"] if live[i] else 0 for i, g in enumerate(games)], dtype=torch.int32, device=vec.device)
            obs, agent_r, _, opp_r, done, info = vec.dual_step(acts, opponent, reward_mode=mode)
            o, m, rows = obs.cpu().numpy(), info["action_mask"].cpu().numpy(), vec.export_state().cpu().numpy()
            for i, g in enumerate(games):
                if not live[i]:
                    continue
                want = g["steps"][t]
                where = (mode, g["wrapper"], g["seed"], t)
                assert float(agent_r[i]) == pytest.approx(want["r"]) and bool(done[i]) == want["done"], where
                assert _sha12(o[i], np.int32) == want["obs"] and _sha12(m[i], np.int8) == want["mask"] and _sha12(rows[i], np.int32) == want["row"], where
                if want.get("opponent_action") is not None:
                    assert int(moves["opp"][i]) == want["opponent_action"] and float(opp_r[i]) == pytest.approx(want["opponent_reward"]), where


def test_random_opponent_helper():
    from splendor_gym_b200.wrappers import random_opponent

    m = np.zeros(45, np.int8)
    assert random_opponent(None, {"action_mask": m}) == 0 and random_opponent(None, {}) == 0
    m[[3, 17]] = 1
    assert {random_opponent(None, {"action_mask": m}) for _ in range(50)} == {3, 17}


# ----------------------------------------------------------------------------- scripts (callers of the hot path)
def test_random_rollout_script(capsys):
    """scripts/random_rollout.py (BASELINE config 1) on the facade and in its batched form."""
    from splendor_gym_b200.scripts import random_rollout

    np.random.seed(0)
    random_rollout.main(["--episodes", "2", "--seed", "0"])
    out = capsys.readouterr().out
    assert "Episode 0: steps=" in out and "Wins:" in out
    st = random_rollout.main(["--episodes", "1", "--envs", "512"])
    assert st["episodes"] >= 512 and st["p0_wins"] + st["p1_wins"] + st["tie_draws"] + st["limit_draws"] + st["nolegal_draws"] == st["episodes"]


def test_ppo_rollout_collection_small():
    """Rollout collection with the MLP policy in the loop (BASELINE config 3, tiny): buffers are filled, actions
    are legal under the stored masks, agent is always player 0, rewards only at episode ends."""
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200.scripts.ppo_rollout import ActorCritic, collect

    torch.manual_seed(0)
    net = ActorCritic().cuda().eval()
    env = SplendorVecEnv(2048, seed=4, shuffle="philox", autoreset=True)
    env.reset()
    buf = collect(env, net, 96)
    legal = buf["masks"].gather(2, buf["actions"].long().unsqueeze(2)).squeeze(2)
    has_move = buf["masks"].sum(dim=2) > 0
    assert bool((legal[has_move] == 1).all())
    assert bool((buf["obs"][:, :, 294] == 0).all())  # the agent always observes as player 0
    r, d = buf["rewards"], buf["terminals"]
    assert bool((r[~d] == 0).all()) and int(d.sum()) > 100
    vals = torch.unique(r[d]).cpu().tolist()
    assert all(min(abs(v - c) for c in (-1.0, 0.0, 1.0, -0.1)) < 1e-6 for v in vals)
    assert int(env.stats[0]) == int(d.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("n,obs_dtype,shuffle", [(1000, "int32", "philox"), (8192 + 32, "int32", "philox"), (4096, "uint8", "philox"),
                                                 (7, "int32", "philox"), (2048, "int32", "mt19937")])
def test_step_host_equals_step(n, obs_dtype, shuffle):
    """The host-buffer path (spl_host_step: compact device outputs, chunked D2H, host widening) returns exactly what
    step() returns on the device: observation, mask, reward, terminated, info bits, sampled next action, statistics."""
    import torch
    from splendor_gym_b200 import SplendorVecEnv

    dt = getattr(torch, obs_dtype)
    a = SplendorVecEnv(n, seed=5, shuffle=shuffle, autoreset=True)
    b = SplendorVecEnv(n, seed=5, shuffle=shuffle, autoreset=True)
    oa, ia = a.reset()
    ob, ib = b.reset_host(obs_dtype=dt, sample_next=True)
    assert ob.device.type == "cpu" and ob.dtype == dt
    assert torch.equal(oa.cpu().to(dt), ob) and torch.equal(ia["action_mask"].cpu(), ib["action_mask"])
    act = a.sample_random_actions().clone()
    assert torch.equal(act.cpu(), b._host["next_action"])
    assert b.host_stats()["gpu_writable"] == 1.0  # the result arrays come from spl_host_alloc
    rng = torch.Generator().manual_seed(0)
    for t in range(70 if shuffle == "philox" else 200):
        if t % 9 == 4:  # sprinkle illegal / out-of-range actions
            bad = torch.randint(0, n, (max(1, n // 50),), generator=rng)
            act[bad.to(act.device)] = torch.randint(-3, 60, (bad.numel(),), generator=rng, dtype=torch.int32).to(act.device)
        h_act = act.cpu().numpy().copy()
        o1, r1, t1, _, i1 = a.step(act, sample_next=True)
        o2, r2, t2, tr2, i2 = b.step_host(h_act, obs_dtype=dt, sample_next=True)
        assert torch.equal(o1.cpu().to(dt), o2), f"step {t}: obs"
        assert torch.equal(i1["action_mask"].cpu(), i2["action_mask"]), f"step {t}: mask"
        assert torch.equal(r1.cpu(), r2) and torch.equal(t1.cpu(), t2) and not tr2.any()
        assert torch.equal(a.info_bits.cpu(), i2["info_bits"])
        assert torch.equal(a.next_action.cpu(), i2["next_action"])
        act = a.next_action.clone()
    assert torch.equal(a.export_state(), b.export_state()) and torch.equal(a.stats, b.stats)
    assert int(a.stats[0]) > 0 or n < 100
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("direct,nibbles,threads", [("0", "1", 3), ("0.5", "1", 2), ("1", "1", 1), ("0.25", "0", 5), ("0", "0", 1)])
def test_step_host_shares_and_transfer_forms(direct, nibbles, threads, monkeypatch):
    """Whatever the split between the GPU-written (already widened) share and the share host threads widen, and whether
    the staged observation travels nibble-packed or raw, the host arrays hold exactly what step() returns.  Includes a
    partial last group (n % 64 != 0) and hand-built states whose token counts do not fit a nibble (raw fallback per group)."""
    import torch
    from splendor_gym_b200 import SplendorVecEnv

    monkeypatch.setenv("SPL_HOST_DIRECT", direct)
    monkeypatch.setenv("SPL_HOST_NIBBLES", nibbles)
    n = 64 * 37 + 21
    a = SplendorVecEnv(n, seed=9, shuffle="philox", autoreset=True)
    b = SplendorVecEnv(n, seed=9, shuffle="philox", autoreset=True)
    b.lib.spl_host_set_threads(threads)
    a.reset()
    b.reset()
    rows = a.export_state()
    big = torch.arange(0, n, 97, device=rows.device)
    rows[big, 0] = 20          # bank white = 20: observation entry 0 does not fit a nibble
    rows[big[::2], 6] = 17     # and 17 white tokens in the mover's hand
    a.import_state(rows)
    b.import_state(rows)
    act = a.sample_random_actions().clone()
    for t in range(12):
        o1, r1, t1, _, i1 = a.step(act, sample_next=True)
        o2, r2, t2, tr2, i2 = b.step_host(act.cpu().numpy().copy(), sample_next=True)
        assert torch.equal(o1.cpu(), o2), f"step {t}: obs"
        assert torch.equal(i1["action_mask"].cpu(), i2["action_mask"]), f"step {t}: mask"
        assert torch.equal(r1.cpu(), r2) and torch.equal(t1.cpu(), t2)
        assert torch.equal(a.info_bits.cpu(), i2["info_bits"]) and torch.equal(a.next_action.cpu(), i2["next_action"])
        st = b.host_stats()
        assert abs(st["gpu_written_share"] - float(direct)) < 0.02 and st["threads"] <= threads
        act = a.next_action.clone()
    assert (o2[:, 0] >= 16).any(), "the hand-built wide entries should still be on the table"
    b.lib.spl_host_set_threads(0)
    b.close()


@pytest.mark.gpu
def test_step_host_into_plain_arrays():
    """Result arrays that the GPU cannot write (ordinary pageable memory handed to the C ABI directly): everything is
    widened by host threads, same values."""
    import ctypes as C
    import numpy as np
    import torch
    from splendor_gym_b200 import SplendorVecEnv, _lib as L

    n = 3000
    a = SplendorVecEnv(n, seed=2, shuffle="philox", autoreset=True)
    b = SplendorVecEnv(n, seed=2, shuffle="philox", autoreset=True)
    a.reset()
    b.reset()
    handle = C.c_void_p()
    L.check(b.lib.spl_host_create(n, 0, C.byref(handle)))
    obs = np.zeros((n, 297), np.int32)
    mask = np.zeros((n, 45), np.int8)
    rew = np.zeros(n, np.float32)
    term = np.zeros(n, np.uint8)
    nxt = np.zeros(n, np.int32)
    io = L.SplHostIO()
    io.obs, io.mask, io.reward, io.terminated, io.next_action = obs.ctypes.data, mask.ctypes.data, rew.ctypes.data, term.ctypes.data, nxt.ctypes.data
    io.action_key, io.autoreset = b.action_key, 1
    act = a.sample_random_actions().clone()
    for t in range(10):
        h_act = act.cpu().numpy().copy()
        io.actions, io.action_t = h_act.ctypes.data, t + 1
        o1, r1, t1, _, i1 = a.step(act, sample_next=True)
        L.check(b.lib.spl_host_step(handle, C.byref(b._envs), C.byref(io), b._stream()))
        assert np.array_equal(o1.cpu().numpy(), obs) and np.array_equal(i1["action_mask"].cpu().numpy(), mask)
        assert np.array_equal(r1.cpu().numpy(), rew) and np.array_equal(t1.cpu().numpy().astype(np.uint8), term)
        assert np.array_equal(a.next_action.cpu().numpy(), nxt)
        st = (C.c_double * 8)()
        b.lib.spl_host_get_stats(handle, st)
        assert st[7] == 0.0 and st[5] == 0.0
        act = a.next_action.clone()
    b.lib.spl_host_destroy(handle)


@pytest.mark.gpu
def test_step_host_rejects_what_it_cannot_do():
    import numpy as np
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200._lib import SplendorB200Error

    e = SplendorVecEnv(64, seed=1, shuffle="mt19937", autoreset=True, prefetch_deals=False)
    e.reset()
    with pytest.raises(SplendorB200Error):
        e.step_host(np.zeros(64, np.int32))
    o, r, t, tr, info = e.step_host(np.zeros(64, np.int32), autoreset=False)  # MT19937 decks are fine without auto-reset
    assert o.shape == (64, 297)
    with pytest.raises(ValueError):
        e.step_host(np.zeros(3, np.int32), autoreset=False)


@pytest.mark.gpu
@pytest.mark.parametrize("n,shuffle", [(33, "philox"), (1000, "philox"), (4096, "philox"), (1000, "mt19937")])
def test_f16_observation_mode_equals_int32(n, shuffle):
    """obs_format='f16' (fp16 [N,304] policy input + uint8 observation, cast fused into the step kernel) carries exactly
    the values of the reference-typed int32 observation; masks, rewards, terminations, info and state are unchanged."""
    import torch
    from splendor_gym_b200 import SplendorVecEnv

    a = SplendorVecEnv(n, seed=9, shuffle=shuffle, autoreset=True)
    b = SplendorVecEnv(n, seed=9, shuffle=shuffle, autoreset=True, obs_format="f16")
    oa, _ = a.reset()
    ob, _ = b.reset()

    def check(t):
        assert b.obs.dtype == torch.uint8 and b.obs_f16.dtype == torch.float16 and b.obs_f16.shape == (n, 304)
        assert torch.equal(a.obs, b.obs.to(torch.int32)), f"step {t}: uint8 obs"
        assert torch.equal(a.obs.to(torch.float16), b.obs_f16[:, :297]), f"step {t}: fp16 obs"
        assert not b.obs_f16[:, 297:].any(), f"step {t}: padding columns"
        assert torch.equal(a.mask, b.mask)

    check(-1)
    act = a.sample_random_actions().clone()
    for t in range(90):
        if t % 11 == 5:
            act[::37] = 44 - act[::37]  # mostly illegal
        a.step(act, sample_next=True)
        b.step(act, sample_next=True)
        check(t)
        assert torch.equal(a.reward, b.reward) and torch.equal(a._terminated, b._terminated) and torch.equal(a.info_bits, b.info_bits)
        assert torch.equal(a.next_action, b.next_action)
        act = a.next_action.clone()
    assert torch.equal(a.export_state(), b.export_state()) and torch.equal(a.stats, b.stats)
    # redirected outputs (rollout-buffer slices) and observe()
    buf8 = torch.zeros((n, 297), dtype=torch.uint8, device="cuda")
    buf16 = torch.zeros((n, 304), dtype=torch.float16, device="cuda")
    a.step(act)
    b.step(act, out_obs=buf8, out_obs_f16=buf16)
    assert torch.equal(a.obs, buf8.to(torch.int32)) and torch.equal(a.obs.to(torch.float16), buf16[:, :297])
    b.observe()
    assert torch.equal(a.obs, b.obs.to(torch.int32)) and torch.equal(a.obs.to(torch.float16), b.obs_f16[:, :297])


@pytest.mark.gpu
def test_f16_mode_argument_checks():
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200._lib import SplendorB200Error

    with pytest.raises(SplendorB200Error):
        SplendorVecEnv(64, shuffle="mt19937", autoreset=True, obs_format="f16", prefetch_deals=False)
    with pytest.raises(ValueError):
        SplendorVecEnv(64, obs_format="bf16")
    e = SplendorVecEnv(64, seed=3, shuffle="mt19937", autoreset=False, obs_format="f16")  # fine without auto-reset
    e.reset()
    e.step(e.sample_random_actions())
    assert e.obs_f16[:, :297].max() < 256


@pytest.mark.gpu
def test_ppo_rollout_f16_policy_input_matches_cast_path():
    """The fp16 policy input emitted by the step kernel + zero-padded first / last layers give the same logits as the reference's
    cast path (obs.to(fp16) through the unpadded network) -- same fp16 GEMM inputs, so the same action choices."""
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200.scripts.ppo_rollout import ActorCritic, collect, pad_first_layer, pad_head

    torch.manual_seed(1)
    net = ActorCritic().cuda().half().eval()
    a = SplendorVecEnv(4096, seed=6, shuffle="philox", autoreset=True)
    b = SplendorVecEnv(4096, seed=6, shuffle="philox", autoreset=True, obs_format="f16")
    a.reset()
    b.reset()
    padded = pad_head(pad_first_layer(net.actor))
    with torch.no_grad():
        ref = net.actor(a.obs.to(torch.float16)).float()
        got = padded(b.obs_f16)[:, :45].float()
    assert torch.allclose(ref, got, atol=2e-3, rtol=2e-3)
    net.actor, net.critic = padded, pad_first_layer(net.critic)
    buf = collect(b, net, 48, dtype=torch.float16)
    assert buf["obs"].dtype == torch.uint8
    legal = buf["masks"].gather(2, buf["actions"].long().unsqueeze(2)).squeeze(2)
    assert bool((legal[buf["masks"].sum(dim=2) > 0] == 1).all()) and int(buf["terminals"].sum()) > 50


@pytest.mark.gpu
def test_import_state_rejects_rows_outside_the_domain():
    """spl_import_state validates instead of truncating: a row with a counter >= 128 or a card id >= 90 is not imported (its env
    keeps its state), the valid rows of the same call are, and the call reports SPL_E_BADROW."""
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200._lib import SplendorB200Error

    env = SplendorVecEnv(8, seed=4, shuffle="mt19937", autoreset=False)
    env.reset()
    before = env.export_state().clone()
    rows = before.clone()
    rows[1, 0] = 200      # bank white = 200 would not fit the byte-wise arithmetic
    rows[5, 52] = 97      # board slot holds a card id outside the 90-card table
    rows[2, 0] = 3        # a valid edit
    with pytest.raises(SplendorB200Error, match="outside the engine's domain"):
        env.import_state(rows)
    after = env.export_state()
    assert torch.equal(after[1], before[1]) and torch.equal(after[5], before[5])
    assert int(after[2, 0]) == 3 and torch.equal(after[[0, 3, 4, 6, 7]], before[[0, 3, 4, 6, 7]])
