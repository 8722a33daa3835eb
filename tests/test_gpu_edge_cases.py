"""GPU edge cases of the step/reset path through the C ABI: a single env (BASELINE config 1), env counts that do
not fill a warp / block, masked resets, TimeLimit truncation, auto_reset off, action checking, checkpoint
round trips.  Same parity bar as test_gpu_parity.py (1e-9 * max(1, |ref|), flags and integer state exact)."""
import numpy as np
import pytest
import torch

import spacegym_b200 as sg
from golden_cases import close
from helpers import compare_engine_to_oracle, oracle_cfg, oracle_state_to_engine, random_actions, to_dev

pytestmark = pytest.mark.gpu


def _lockstep(oracle, case, env_id, n, T, seed, mode, **kw):
    cfg = oracle_cfg(oracle, case, **{k: v for k, v in kw.items() if k == "max_episode_steps"})
    eng = oracle.Engine(cfg, n, seed=seed)
    env = sg.make(env_id, n, seed=seed, step_mode=mode, **kw)
    assert close(env.reset().cpu().numpy(), eng.reset()).all()
    rng = np.random.RandomState(seed)
    bad = 0
    for t in range(T):
        a = random_actions(rng, 1, n, bool(cfg.discrete))
        o_o, r_o, d_o, tr_o = eng.step(a)
        o_d, r_d, d_d, info = env.step(to_dev(a[0]))
        ok = close(o_d.cpu().numpy(), o_o).all(axis=1) & close(r_d.cpu().numpy(), r_o[0])
        ok &= (d_d.cpu().numpy() == d_o[0].astype(bool))
        ok &= info["TimeLimit.truncated"].cpu().numpy() == tr_o[0].astype(bool)
        ok &= compare_engine_to_oracle(cfg, env.get_state(), eng.envs)
        if (~ok).any():
            bad += int((~ok).sum())
            env.set_state(oracle_state_to_engine(cfg, eng.envs))
    env.close()
    return bad, eng


def test_config1_single_env_1000_steps(oracle):
    """BASELINE config 1: GoalContinuous2P-v0, one env, 1000 random-action steps (20+ episodes with auto-reset)."""
    bad, eng = _lockstep(oracle, "goal2p", "GoalContinuous2P-v0", 1, 1000, seed=5, mode="split")
    assert bad == 0 and eng.stats[2] >= 5


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 257])
@pytest.mark.parametrize("mode", ["split", "fused"])
def test_ragged_env_counts(oracle, n, mode):
    """Env counts that leave partial warps / blocks (the reference has no batch at all: any N must work)."""
    bad, _ = _lockstep(oracle, "goal3p", "GoalContinuous3P-v0", n, 80, seed=100 + n, mode=mode, max_episode_steps=30)
    assert bad == 0
    bad, _ = _lockstep(oracle, "donotcrash", "DoNotCrashContinuous-v0", n, 40, seed=200 + n, mode=mode,
                       max_episode_steps=30)
    assert bad == 0


def test_masked_reset_only_touches_selected_envs():
    n = 300
    env = sg.make("GoalContinuous3P-v0", n, seed=9)
    env.reset()
    env.step(to_dev(np.random.RandomState(0).uniform(-1, 1, size=(n, 2)).astype(np.float32)))
    before = env.get_state()
    mask = torch.zeros(n, dtype=torch.uint8)
    mask[::7] = 1
    obs = env.reset(mask)
    after = env.get_state()
    m = mask.bool().numpy()
    for k in ("ship", "planets", "goal", "elapsed", "ep_return", "tile_word"):
        a, b = after[k].cpu().numpy().reshape(n, -1), before[k].cpu().numpy().reshape(n, -1)
        assert (a[~m] == b[~m]).all(), k
    assert (after["elapsed"].cpu().numpy()[m] == 0).all() and (before["elapsed"].cpu().numpy()[~m] == 1).all()
    assert (after["ship"].cpu().numpy()[m] != before["ship"].cpu().numpy()[m]).any(axis=1).all()
    # the returned observation rows of the reset envs describe the new state (x, y in slots 0, 1)
    assert close(obs.cpu().numpy()[m][:, :2], after["ship"].cpu().numpy()[m][:, :2]).all()
    env.close()


def test_time_limit_truncates_like_gym(oracle):
    """gym 0.21 TimeLimit: at elapsed == max_episode_steps, done = True and truncated = not terminated."""
    n, L = 512, 7
    env = sg.make("KeplerCircleOrbit-v0", n, seed=1, max_episode_steps=L)
    env.reset()
    zeros = torch.zeros((n, 2), dtype=torch.float32, device=env.device)
    for t in range(1, 2 * L + 1):
        _, _, done, info = env.step(zeros)
        tr = info["TimeLimit.truncated"]
        el = env.get_state()["elapsed"].cpu().numpy()
        assert (tr & ~done).sum() == 0  # truncated implies done
        if t % L == 0:  # nothing terminates within 14 coasting steps from these starts: every env is truncated here
            assert done.all() and tr.all() and (el == 0).all()
        else:
            assert not done.any() and (el == t % L).all()
    assert env.stats()["episodes"] == 2 * n and env.stats()["length_sum"] == 2 * n * L
    env.close()


def test_auto_reset_off_keeps_terminal_state(oracle):
    """auto_reset=0: a done env is not restarted by step (the caller resets it, as with a bare gym.Env)."""
    n = 2048
    cfg = oracle_cfg(oracle, "goal3p")
    env = sg.make("GoalContinuous3P-v0", n, seed=4, auto_reset=0, max_episode_steps=0)
    eng = oracle.Engine(cfg, n, seed=4)
    env.reset()
    eng.reset()
    rng = np.random.RandomState(1)
    seen = np.zeros(n, dtype=bool)
    for t in range(60):
        a = random_actions(rng, 1, n, False)
        o_o, r_o, d_o, _ = eng.step(a, auto_reset=False)
        o_d, r_d, d_d, _ = env.step(to_dev(a[0]))
        d = d_d.cpu().numpy()
        fresh = ~seen  # compare until an env's first termination (afterwards the two drift by design of the test)
        assert (d[fresh] == d_o[0].astype(bool)[fresh]).all()
        assert close(o_d.cpu().numpy()[fresh], o_o[fresh]).all() and close(r_d.cpu().numpy()[fresh], r_o[0][fresh]).all()
        seen |= d
    assert seen.sum() > 50
    assert (env.get_state()["elapsed"].cpu().numpy()[seen] > 0).all()  # never restarted
    env.close()


def test_action_checking_matches_reference_asserts():
    """spaceship_env.py:71 asserts action_space.contains(action); opt-in here (it costs a device sync)."""
    env = sg.make("GoalContinuous3P-v0", 64, seed=0)
    env.reset()
    bad = torch.full((64, 2), 1.5, dtype=torch.float32, device=env.device)
    with pytest.raises(AssertionError):
        env.step(bad, check_actions=True)
    with pytest.raises(AssertionError):
        env.step(torch.zeros((63, 2), dtype=torch.float32, device=env.device))  # wrong batch shape
    with pytest.raises(AssertionError):
        env.step(torch.zeros((64, 2), dtype=torch.float64, device=env.device))  # wrong dtype
    env.close()
    denv = sg.make("DoNotCrashDiscrete-v0", 64, seed=0)
    denv.reset()
    a = torch.zeros((64, 2), dtype=torch.float32, device=denv.device)
    a[:, 0] = 6
    with pytest.raises(AssertionError):
        denv.step(a, check_actions=True)
    denv.close()


def test_checkpoint_round_trip_resumes_bit_identically():
    """get_state -> set_state into a fresh handle (same seed): both continue with identical outputs, including
    auto-resets and goal resampling (the per-env Philox counters are part of the state)."""
    n = 4096
    a = sg.make("GoalContinuous4P-v0", n, seed=21, max_episode_steps=40)
    a.reset()
    a.rollout_random(50)
    snap = a.get_state()
    b = sg.make("GoalContinuous4P-v0", n, seed=21, max_episode_steps=40)
    b.reset()
    b.set_state(snap)
    for _ in range(3):
        oa, ra, da, _ = a.step_many(K=20)
        ob, rb, db, _ = b.step_many(K=20)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert (sa[k] == sb[k]).all() if k != "step_count" else sa[k] == sb[k], k
    a.close()
    b.close()


@pytest.mark.parametrize("env_id", ["GoalContinuous2P-v0", "GoalContinuous3P-v0", "KeplerEllipseEasy-v0",
                                    "DoNotCrashContinuous-v0"])
@pytest.mark.parametrize("n", [1000, 4097])
def test_float32_observations_every_layout(env_id, n):
    """obs_dtype=float32 through the split path (K = 1, staged rows + 128-bit warp stores) for every observation
    width (13, 15, 10, 7) and env counts that end in a partial warp: equals the cast of the float64 rows."""
    a = sg.make(env_id, n, seed=6, max_episode_steps=25)
    b = sg.make(env_id, n, seed=6, max_episode_steps=25, obs_dtype="float32")
    oa, ob = a.reset(), b.reset()
    assert ob.dtype == torch.float32 and torch.equal(oa.float(), ob)
    g = torch.Generator(device=a.device).manual_seed(1)
    for _ in range(40):
        act = (torch.rand((n, 2), device=a.device, generator=g) * 2 - 1).float()
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert torch.equal(oa.float(), ob) and torch.equal(ra, rb) and torch.equal(da, db)
    a.close()
    b.close()


@pytest.mark.parametrize("mode", ["split", "fused"])
def test_non_finite_state_terminates_and_is_flagged(mode):
    """A NaN / Inf state injected through set_state must not hang the integrator loops: the env is flagged
    (status bits: step too small / non-finite) like the reference's `assert ode_solution.success`
    (dynamic_model.py:120), and every other env of the batch steps on unharmed."""
    n = 512
    env = sg.make("GoalContinuous3P-v0", n, seed=3, step_mode=mode, auto_reset=0, max_episode_steps=0)
    ref = sg.make("GoalContinuous3P-v0", n, seed=3, step_mode=mode, auto_reset=0, max_episode_steps=0)
    env.reset()
    ref.reset()
    st = env.get_state()
    ship = st["ship"].clone()
    ship[5, 0] = float("nan")      # x
    ship[77, 3] = float("inf")     # vx
    ship[300, 2] = float("nan")    # heading
    env.set_state({"ship": ship})
    act = torch.zeros((n, 2), dtype=torch.float32, device=env.device)
    for _ in range(3):
        o, r, d, _ = env.step(act)
        o2, r2, d2, _ = ref.step(act)
    torch.cuda.synchronize()
    status = env.get_state()["status"].cpu().numpy()
    bad = np.zeros(n, dtype=bool)
    bad[[5, 77, 300]] = True
    assert (status[bad] != 0).all() and (status[~bad] == 0).all()
    assert torch.equal(o[~torch.as_tensor(bad, device=o.device)], o2[~torch.as_tensor(bad, device=o.device)])
    env.close()
    ref.close()


def test_single_env_adapter_runs_the_reference_loop(oracle):
    """SpaceEnv: the reference's own call pattern (gym.make -> seed -> reset -> step ... if done: reset) for ONE env,
    numpy in / numpy and python scalars out, TimeLimit info key, action asserts, rgb_array render — in lockstep with the
    oracle twin stepped without auto-reset (BASELINE configs[0] shape: GoalContinuous2P-v0, one env)."""
    cfg = oracle_cfg(oracle, "goal2p", max_episode_steps=40)
    eng = oracle.Engine(cfg, 1, seed=3)
    env = sg.make_single("GoalContinuous2P-v0", seed=11, max_episode_steps=40)
    with pytest.raises(AssertionError):
        env.step(np.zeros(2, dtype=np.float32))  # before the first reset (gym TimeLimit)
    assert env.seed(3) == [3]
    assert env.observation_space.shape == (13,) and env.action_space.shape == (2,)
    obs = env.reset()
    assert obs.shape == (13,) and close(obs, eng.reset()[0]).all()
    rng = np.random.RandomState(8)
    episodes = truncated = 0
    for t in range(400):
        a = rng.uniform(-1, 1, size=2).astype(np.float32)
        o, r, d, info = env.step(a)
        o_o, r_o, d_o, tr_o = eng.step(a[None, None], auto_reset=False)
        assert isinstance(r, float) and isinstance(d, bool) and o.shape == (13,)
        assert close(o, o_o[0]).all() and close(r, r_o[0][0]) and d == bool(d_o[0][0])
        assert ("TimeLimit.truncated" in info) == bool(tr_o[0][0])
        truncated += int("TimeLimit.truncated" in info)
        if d:
            episodes += 1
            assert close(env.reset(), eng.reset_live()[0]).all()
    assert episodes >= 8 and truncated >= 1
    with pytest.raises(AssertionError):
        env.step(np.array([2.0, 0.0], dtype=np.float32))  # outside Box(-1, 1): spaceship_env.py:71
    frame = env.render()
    assert frame.ndim == 3 and frame.shape[2] == 3 and frame.dtype == np.uint8 and frame.any()
    assert close(env.vector_field(np.zeros(2, dtype=np.float32))[:2], env.get_state()["ship"][0, 3:5].cpu().numpy()).all()
    env.close()


@pytest.mark.parametrize("case,env_id", [("goal3p", "GoalContinuous3P-v0"), ("kepler_random", "KeplerRandomOrbits-v0")])
def test_masked_reset_of_a_live_batch_matches_the_oracle(oracle, case, env_id):
    """sg_reset on a running batch (the reference's `if done: env.reset()` for a subset of envs): the masked envs restart
    from the next block of their own stream, the others keep their state; observation rows and the whole state against
    the oracle twin's live reset, then stepping goes on in lockstep."""
    n, seed = 1000, 21
    cfg = oracle_cfg(oracle, case)
    eng = oracle.Engine(cfg, n, seed=seed)
    from helpers import engine_kwargs
    _, kw = engine_kwargs(case)
    env = sg.make(env_id, n, seed=seed, **kw)
    assert close(env.reset().cpu().numpy(), eng.reset()).all()
    rng = np.random.RandomState(4)
    for rep in range(3):
        for _ in range(5):
            a = random_actions(rng, 1, n, bool(cfg.discrete))
            eng.step(a)
            env.step(to_dev(a[0]))
        mask = (rng.uniform(size=n) < 0.3).astype(np.uint8)
        obs_d = env.reset(torch.as_tensor(mask))  # a CPU mask is moved to the device by the host layer
        obs_o = eng.reset_live(mask)
        assert close(obs_d.cpu().numpy(), obs_o).all()
        assert compare_engine_to_oracle(cfg, env.get_state(), eng.envs).all()
    a = random_actions(rng, 1, n, bool(cfg.discrete))
    o_o, r_o, d_o, _ = eng.step(a)
    o_d, r_d, d_d, _ = env.step(to_dev(a[0]))
    assert close(o_d.cpu().numpy(), o_o).all() and close(r_d.cpu().numpy(), r_o[0]).all()
    env.close()


@pytest.mark.parametrize("case,env_id", [("goal3p", "GoalContinuous3P-v0"), ("kepler_circle", "KeplerCircleOrbit-v0"),
                                          ("goal3p_discrete", None)])
def test_terminal_observation_under_auto_reset(oracle, case, env_id):
    """sg_set_terminal_obs / info["terminal_observation"]: for every env that restarts in a step (terminal event or
    TimeLimit) the observation of the state its episode ended in — what the reference's loop sees with done = True
    (spaceship_env.py:68-78) — against the oracle twin; rows of envs that did not finish keep their old content."""
    from helpers import engine_kwargs
    n, seed = 3000, 6
    cfg = oracle_cfg(oracle, case, max_episode_steps=25)
    eng = oracle.Engine(cfg, n, seed=seed)
    eid, kw = engine_kwargs(case)
    env = sg.make(eid, n, seed=seed, max_episode_steps=25, **kw)
    term_d = env.enable_terminal_observation()
    term_o = np.zeros((n, env.obs_dim))
    assert close(env.reset().cpu().numpy(), eng.reset()).all()
    rng = np.random.RandomState(12)
    by_event = by_limit = 0
    for t in range(60):
        a = random_actions(rng, 1, n, bool(cfg.discrete))
        o_o, r_o, d_o, tr_o = eng.step(a, term_obs=term_o)
        o_d, r_d, d_d, info = env.step(to_dev(a[0]))
        d = d_d.cpu().numpy()
        assert (d == d_o[0].astype(bool)).all() and close(o_d.cpu().numpy(), o_o).all()
        t_d = info["terminal_observation"].cpu().numpy()
        assert close(t_d, term_o).all(), f"step {t}: terminal observation rows differ"
        tr = info["TimeLimit.truncated"].cpu().numpy()
        by_limit += int(tr.sum())
        by_event += int((d & ~tr).sum())
        # a finished env's terminal row is NOT the (reset) observation returned by the step
        if d.any():
            assert (np.abs(t_d[d] - o_d.cpu().numpy()[d]).max(axis=1) > 0).all()
    assert by_event >= 1 and by_limit > 1000  # both kinds of endings were seen (Kepler orbits rarely crash in 25 steps)
    assert info["terminal_observation"].data_ptr() != term_d.data_ptr()  # copy=True hands out a fresh tensor
    env.enable_terminal_observation(False)
    _, _, _, info = env.step(to_dev(random_actions(rng, 1, n, bool(cfg.discrete))[0]))
    assert "terminal_observation" not in info
    env.close()
