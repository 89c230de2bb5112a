"""Device-side callers of the step path: scripted opponents vs the decisions the
reference's bots made on the same (obs, mask) (tests/golden/bots.json), masked categorical sampling vs
torch.distributions, GAE vs a float64 restatement of ppo_splendor.py:307-314."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_scripted_bots_match_reference_decisions():
    from splendor_gym_b200.policy import scripted_action

    recs = load_golden("bots.json")
    obs = torch.tensor([r["obs"] for r in recs], dtype=torch.int32, device="cuda")
    mask = torch.tensor([r["mask"] for r in recs], dtype=torch.int8, device="cuda")
    g1 = scripted_action(obs, mask, "greedy_v1").cpu().tolist()
    g2 = scripted_action(obs, mask, "greedy_v2").cpu().tolist()
    assert g1 == [r["greedy_v1"] for r in recs]
    assert g2 == [r["greedy_v2"] for r in recs]
    seen = [set() for _ in recs]
    for t in range(48):
        b = scripted_action(obs, mask, "basic", t=t).cpu().tolist()
        rnd = scripted_action(obs, mask, "random", t=t).cpu().tolist()
        for i, r in enumerate(recs):
            assert b[i] in r["basic_support"], (i, b[i], r["basic_support"])
            assert r["mask"][rnd[i]] == 1 or sum(r["mask"]) == 0
            seen[i].add(b[i])
    # the random tie-breaks cover the whole support the reference's np.random.choice can produce
    assert all(seen[i] == set(r["basic_support"]) for i, r in enumerate(recs) if len(r["basic_support"]) <= 4)


def test_bots_drive_dual_step():
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200.policy import bot_policy, scripted_action

    env = SplendorVecEnv(4096, seed=12, shuffle="philox", autoreset=True)
    env.reset()
    opp = bot_policy("greedy_v1")
    for t in range(120):
        a = scripted_action(env.obs, env.mask, "basic", t=t)
        env.dual_step(a, opp)
    st = env.stats.cpu().tolist()
    assert st[0] > 2000 and st[1] + st[2] + st[3] + st[4] + st[5] == st[0]


def test_masked_sample_f16_pitched_equals_float_path():
    """fp16 logits read in place from a padded [N,48] head (spl_masked_sample_f16) give exactly the results of the float path
    on the same values: actions, log-probs, entropy."""
    from splendor_gym_b200.policy import masked_sample

    torch.manual_seed(2)
    n = 5000
    head = (torch.randn(n, 48, device="cuda") * 2).half()
    mask = (torch.rand(n, 45, device="cuda") < 0.3).to(torch.int8)
    mask[:5] = 0
    view = head[:, :45]
    assert not view.is_contiguous()
    for greedy in (False, True):
        a1, lp1, e1 = masked_sample(view, mask, greedy=greedy, t=4, want_entropy=True)
        a2, lp2, e2 = masked_sample(view.float().contiguous(), mask, greedy=greedy, t=4, want_entropy=True)
        assert torch.equal(a1, a2) and torch.equal(lp1, lp2) and torch.equal(e1, e2)
    dense = head[:, :45].contiguous()  # pitch 45, unaligned rows
    a3, lp3, _ = masked_sample(dense, mask, t=4)
    a4, lp4, _ = masked_sample(dense.float(), mask, t=4)
    assert torch.equal(a3, a4) and torch.equal(lp3, lp4)


def test_masked_sample_against_torch():
    from splendor_gym_b200.policy import masked_sample

    torch.manual_seed(0)
    n = 4096
    logits = torch.randn(n, 45, device="cuda") * 2
    mask = (torch.rand(n, 45, device="cuda") < 0.3).to(torch.int8)
    mask[:7] = 0  # rows without a legal action stay unmasked (ppo_splendor.py:36-37)
    m = mask.bool()
    any_legal = m.any(dim=1, keepdim=True)
    ml = torch.where(~m & any_legal, torch.full_like(logits, float("-inf")), logits)
    dist = torch.distributions.Categorical(logits=ml)
    a, lp, ent = masked_sample(logits, mask, greedy=True, want_entropy=True)
    assert torch.equal(a.long(), ml.argmax(dim=1))
    assert torch.allclose(lp, dist.log_prob(a.long()), atol=1e-5, rtol=1e-5)
    assert torch.allclose(ent, dist.entropy(), atol=1e-4, rtol=1e-4)
    a, lp, _ = masked_sample(logits, mask, t=3)
    assert bool((ml.gather(1, a.long().view(-1, 1)) > float("-inf")).all())
    assert torch.allclose(lp, dist.log_prob(a.long()), atol=1e-5, rtol=1e-5)
    # distribution: one fixed row sampled many times with different counters
    row_l = logits[100:101].repeat(20000, 1).contiguous()
    row_m = mask[100:101].repeat(20000, 1).contiguous()
    s, _, _ = masked_sample(row_l, row_m, t=9, want_logprob=False)
    counts = torch.bincount(s.long(), minlength=45).float().cpu().numpy()
    p = dist.probs[100].cpu().numpy()
    exp = p * 20000
    sel = exp > 5
    chi2 = float(((counts[sel] - exp[sel]) ** 2 / exp[sel]).sum())
    assert counts[~(p > 0)].sum() == 0 and chi2 < 4 * sel.sum() + 20


def test_gae_against_reference_formula():
    from splendor_gym_b200.policy import gae

    rng = np.random.RandomState(1)
    T, n, gamma, lam = 128, 3000, 0.99, 0.95
    rewards = (rng.rand(T, n) < 0.02) * rng.choice([-1.0, 1.0, -0.1], size=(T, n))
    terms = rewards != 0
    values = rng.randn(T, n).astype(np.float32)
    last = rng.randn(n).astype(np.float32)
    adv = np.zeros((T, n))
    lastgaelam = np.zeros(n)
    for t in reversed(range(T)):  # ppo_splendor.py:307-313
        nonterm = 1.0 - terms[t]
        nextv = last if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * nextv * nonterm - values[t]
        adv[t] = lastgaelam = delta + gamma * lam * nonterm * lastgaelam
    ret = adv + values
    a, r = gae(torch.tensor(rewards, dtype=torch.float32).cuda(), torch.tensor(values).cuda(), torch.tensor(terms).cuda(),
               torch.tensor(last).cuda(), gamma, lam)
    assert np.allclose(a.cpu().numpy(), adv, atol=2e-5, rtol=1e-5)
    assert np.allclose(r.cpu().numpy(), ret, atol=2e-5, rtol=1e-5)


def test_batched_eval_vs_opponent():
    """eval_vs_opponent (scripts/eval_suite.py:162-208), batched: greedy_v1 beats the random opponent, and playing
    a bot against itself-ish keeps the bookkeeping consistent."""
    from splendor_gym_b200.policy import scripted_action
    from splendor_gym_b200.scripts.eval_suite import eval_vs_opponent

    res = eval_vs_opponent(lambda obs, mask: scripted_action(obs, mask, "greedy_v1"), "random", n_games=2000, seed=1)
    assert res["n"] == 2000 and res["wins"] + res["losses"] + res["draws"] == 2000
    assert res["win_rate"] > 0.6 and res["illegal_action_rate"] < 0.05 and 5 < res["avg_turns"] <= 100
    res2 = eval_vs_opponent(lambda obs, mask: scripted_action(obs, mask, "random", t=7), "basic", n_games=1000, seed=2)
    assert res2["win_rate"] < 0.5


def test_ppo_training_smoke():
    """The end-to-end trainer (device rollout + PPO update of ppo_splendor.py:327-361) runs, fills the pool, evaluates."""
    from splendor_gym_b200.scripts import ppo_train

    log = ppo_train.main(["--num-envs", "512", "--num-steps", "32", "--total-timesteps", str(512 * 32 * 4), "--minibatch-size", "4096",
                          "--snapshot-every-updates", "1", "--eval-every-updates", "4", "--eval-games", "256", "--update-epochs", "2"])
    assert len(log) == 4 and all(r["rollout_agent_steps_per_s"] > 0 for r in log)
    assert "win_rate_vs_random" in log[-1] and 0.0 <= log[-1]["win_rate_vs_random"] <= 1.0
    assert log[-1]["episodes"] > 0


def test_trajectory_log_and_replay():
    """4 bytes per env-step (+ the start state) regenerate the whole rollout: replaying the action log reproduces the
    rewards / terminations / final state of the recorded run and every observation of a directly recorded run."""
    from splendor_gym_b200 import SplendorVecEnv
    from splendor_gym_b200.trajectory import Trajectory, describe_action, record_random, replay

    n, T = 2048, 160
    env = SplendorVecEnv(n, seed=21, shuffle="philox", autoreset=True)
    env.reset()
    for _ in range(10):  # start the log mid-game
        env.step(env.sample_random_actions())
    # reference run that keeps observations
    twin = SplendorVecEnv(n, seed=21, shuffle="philox", autoreset=True)
    twin.reset()
    for _ in range(10):
        twin.step(twin.sample_random_actions())
    traj = record_random(env, T, seed=21)
    assert traj.bytes_per_env_step < 5.5
    env2, rew, term, obs = replay(traj, keep_obs=True)
    assert torch.equal(rew, traj.rewards) and torch.equal(term, traj.terminated)
    assert torch.equal(env2.export_state(), env.export_state())
    for t in range(T):
        o, *_ = twin.step(traj.actions[t])
        assert torch.equal(o, obs[t]), t
    assert int(traj.terminated.sum()) > 1000
    assert describe_action(12, env.export_state()[0].cpu().numpy()) == "Take2: gg"  # (every string: test_logger_strings_and_can_afford...)


@pytest.mark.gpu
def test_opponent_pool_follows_the_reference_sampling_rules():
    """OpponentPool against the rules of ppo_splendor.py:135-143,366-370 and training_utils.py:263-276: current policy while
    the pool is empty, p_current afterwards, snapshots uniform, resampling only where an episode ended, FIFO pool with a
    dropped snapshot kept for the envs still playing against it, greedy masked argmax of the env's OWN opponent."""
    import torch
    from splendor_gym_b200.scripts.ppo_rollout import ActorCritic
    from splendor_gym_b200.scripts.ppo_train import OpponentPool

    torch.manual_seed(3)
    dev = torch.device("cuda")
    n = 60000
    net = ActorCritic().to(dev)
    pool = OpponentPool(net, n, dev, p_current=0.25, pool_size=3)
    everyone = torch.ones(n, dtype=torch.bool, device=dev)
    pool.resample(everyone)
    assert bool((pool.opp_id == -1).all())  # len(pool) == 0 -> the current policy (:139)
    for k in range(3):
        with torch.no_grad():
            for p in net.actor.parameters():
                p.add_(0.05 * torch.randn_like(p))  # every snapshot is a different network
        pool.add_snapshot()
    pool.resample(everyone)
    frac = [(pool.opp_id == k).float().mean().item() for k in (-1, 0, 1, 2)]
    assert abs(frac[0] - 0.25) < 0.01 and all(abs(f - 0.25) < 0.01 for f in frac[1:])  # p_current, then uniform over 3 snapshots
    before = pool.opp_id.clone()
    half = torch.arange(n, device=dev) % 2 == 0
    pool.resample(half)
    assert torch.equal(pool.opp_id[~half], before[~half]) and not torch.equal(pool.opp_id[half], before[half])
    # a fourth snapshot drops id 0 from the pool; envs playing against it keep it until their episode ends
    pool.add_snapshot()
    assert pool.pool == [1, 2, 3] and 0 in pool.models and bool((pool.opp_id == 0).any())
    obs = torch.randint(0, 6, (n, 297), device=dev, dtype=torch.int32)
    mask = (torch.rand(n, 45, device=dev) < 0.3).to(torch.int8)
    mask[:, 0] = 1
    act = pool.act(obs, mask)
    with torch.no_grad():
        for k in (-1, 0, 1, 2):
            sel = pool.opp_id == k
            model = net.actor if k < 0 else pool.models[k]
            logits = model(obs[sel].float()).masked_fill(mask[sel] < 0.5, float("-inf"))  # training_utils.py:270-276
            assert torch.equal(act[sel].long(), logits.argmax(dim=-1)), k
    pool.resample(everyone)
    assert not bool((pool.opp_id == 0).any())
    pool.add_snapshot()
    assert 0 not in pool.models and pool.pool == [2, 3, 4]
