"""GPU parity tests: the CUDA engine (through the C ABI) against the golden vectors of the
reference and against the CPU oracle.  Bar: |delta| <= 1e-9 * max(1, |ref|) in FP64 for
state, observation and reward; done / goal flags and all integer state bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import spacegym_b200 as sg
from golden_cases import CASES, RESET_CASES, TOL, close, load
from helpers import (compare_engine_to_oracle, engine_kwargs, oracle_cfg, oracle_state_to_engine, random_actions,
                     to_dev)

pytestmark = pytest.mark.gpu


def _make(case, n, seed=0, **extra):
    env_id, kw = engine_kwargs(case)
    kw.update(extra)
    return sg.make(env_id, n, seed=seed, **kw)


@pytest.mark.parametrize("mode", ["split", "fused"])
@pytest.mark.parametrize("case", list(CASES))
def test_transitions_match_reference_golden(case, mode):
    """Identical (state, action) pairs harvested from the reference -> same next state, obs, reward, flags."""
    g = load("transitions", case)
    n = len(g["reward"])
    env = _make(case, n, auto_reset=0, max_episode_steps=0, step_mode=mode)
    st = {"ship": g["state0"]}
    if env.cfg.task == 0:
        st["planets"] = g["planets"].reshape(n, -1)
        st["goal"] = g["goal"]
    elif env.cfg.task == 1:
        st["orbit"] = np.stack([g["orbit"][:, 1], g["orbit"][:, 0]], axis=1)  # golden: (angle, ecc, a)
    env.set_state(st)
    obs, rew, done, _ = env.step(to_dev(g["action"]))
    new = env.get_state()
    assert int(new["status"].abs().sum()) == 0
    assert (done.cpu().numpy() == g["done"]).all()
    assert close(new["ship"].cpu().numpy(), g["state1"]).all()
    assert close(obs.cpu().numpy(), g["obs"]).all()
    assert close(rew.cpu().numpy(), g["reward"]).all()
    if env.cfg.task == 0:
        changed = (new["goal"].cpu().numpy() != g["goal"]).any(axis=1)
        assert (changed == g["goal_reached"]).all()
    env.close()


LOCKSTEP = [(c, "split") for c in ("goal2p", "goal3p", "goal4p", "goal3p_accel", "kepler_circle",
                                    "kepler_hard_accel_rand", "goal3p_discrete", "kepler_discrete", "donotcrash",
                                    "donotcrash_vel", "donotcrash_discrete", "goal1p")]
LOCKSTEP += [(c, "fused") for c in ("goal3p", "goal2p", "kepler_hard_accel_rand", "goal3p_discrete", "donotcrash",
                                    "goal1p")]


def _assert_boundary_ties(oracle, cfg, prev_envs, actions, bad_idx):
    """Every env-step on which the engine and the oracle parted must sit on a decision boundary: the oracle's
    margin (min distance of an accept/reject, event-sign, h0-branch or goal-reached decision from its threshold,
    oracle/sg_oracle.h so_step_diag.margin) is at rounding level there.  Anything else is a bug, not a tie."""
    P = int(cfg.n_planets)
    for k in bad_idx:
        e = prev_envs[k]
        planets = np.array(e["planets"][:2 * max(P, 1)]).reshape(-1, 2)
        orbit = (float(e["orbit_angle"]), float(e["orbit_ecc"]), float(cfg.ref_orbit_a))
        r = oracle.step_transition(cfg, np.array(e["y"]), planets, np.array(e["goal"]), orbit, actions[k])
        assert r["margin"] < 1e-12, f"env {k} diverged away from any decision boundary (margin {r['margin']:.3e})"


@pytest.mark.parametrize("case,mode", LOCKSTEP)
def test_rollout_lockstep_with_oracle_engine(oracle, case, mode):
    """Whole engine (auto-reset, TimeLimit, Philox streams, goal resampling) in lockstep with the CPU twin,
    through the split kernels (hot/continuation/tail + work queues) and through the fused kernel."""
    n, T, seed, first = 2048, 160, 1234, 7_000_000_000
    cfg = oracle_cfg(oracle, case, max_episode_steps=60)
    eng = oracle.Engine(cfg, n, seed=seed, first_env_index=first)
    env_id, kw = engine_kwargs(case)
    env = sg.make(env_id, n, seed=seed, first_env_index=first, max_episode_steps=60, step_mode=mode, **kw)
    obs_o = eng.reset()
    obs_d = env.reset()
    assert close(obs_d.cpu().numpy(), obs_o).all()
    assert compare_engine_to_oracle(cfg, env.get_state(), eng.envs).all()
    diverged = _lockstep(oracle, cfg, eng, env, n, T, np.random.RandomState(5))
    assert diverged <= 2, f"{diverged} env-steps diverged out of {n * T}"
    so, sd = eng.stats, env.stats()
    assert abs(sd["episodes"] - so[2]) <= 2 and sd["episodes"] > n
    env.close()


def _lockstep(oracle, cfg, eng, env, n, T, rng):
    """T env steps of the engine and the oracle twin from the same actions; returns the number of env-steps on
    which they parted, each of them proven to be a boundary tie (then the engine is re-synchronised)."""
    diverged = 0
    for t in range(T):
        a = random_actions(rng, 1, n, bool(cfg.discrete))
        prev = eng.envs.copy()
        o_o, r_o, d_o, tr_o = eng.step(a)
        o_d, r_d, d_d, info = env.step(to_dev(a[0]))
        ok = close(o_d.cpu().numpy(), o_o).all(axis=1) & close(r_d.cpu().numpy(), r_o[0])
        ok &= (d_d.cpu().numpy() == d_o[0].astype(bool)) & (info["TimeLimit.truncated"].cpu().numpy() == tr_o[0].astype(bool))
        ok &= compare_engine_to_oracle(cfg, env.get_state(), eng.envs)
        bad = np.flatnonzero(~ok)
        if len(bad):  # boundary ties can flip a discrete decision: prove it, count, re-synchronise from the oracle
            _assert_boundary_ties(oracle, cfg, prev, a[0], bad)
            diverged += len(bad)
            env.set_state(oracle_state_to_engine(cfg, eng.envs))
    return diverged


def test_config2_kepler_circle_4096_envs_lockstep(oracle):
    """BASELINE configs[1] / SURVEY §8d-2 verbatim: KeplerCircleOrbit-v0, 4096 envs, 64 steps in lockstep with the
    oracle twin from identical actions (state / observation / reward to 1e-9 max(1, |ref|), done exact)."""
    n, T, seed = 4096, 64, 0
    cfg = oracle.config_from_id("KeplerCircleOrbit-v0")
    eng = oracle.Engine(cfg, n, seed=seed)
    env = sg.make("KeplerCircleOrbit-v0", n, seed=seed)
    assert close(env.reset().cpu().numpy(), eng.reset()).all()
    diverged = _lockstep(oracle, cfg, eng, env, n, T, np.random.RandomState(2))
    assert diverged <= 1, f"{diverged} env-steps diverged out of {n * T}"
    env.close()


@pytest.mark.parametrize("case", ["goal3p", "kepler_random"])
def test_fused_substeps_and_synthetic_actions(oracle, case):
    """K fused substeps (one launch, state in registers) with in-kernel random actions == K single steps
    through the split kernels == the oracle twin.  The two GPU paths share their device functions but are
    compiled into different kernels (different FMA contraction), so they are compared to the parity bar."""
    n, K, seed = 4096, 12, 99
    cfg = oracle_cfg(oracle, case, max_episode_steps=40)
    env_id, kw = engine_kwargs(case)
    a = sg.make(env_id, n, seed=seed, max_episode_steps=40, step_mode="fused", **kw)
    b = sg.make(env_id, n, seed=seed, max_episode_steps=40, step_mode="split", **kw)
    eng = oracle.Engine(cfg, n, seed=seed)
    a.reset(); b.reset(); eng.reset()
    for rep in range(3):
        oa, ra, da, _ = a.step_many(K=K)
        rb, db = [], []
        for k in range(K):
            ob, r1, d1, _ = b.step_many(K=1)
            rb.append(r1[0]); db.append(d1[0])
        rb, db = torch.stack(rb), torch.stack(db)
        oo, ro, do, _ = eng.step(None, K=K)
        for o_, r_, d_ in ((oa, ra, da), (ob, rb, db)):
            ok = close(o_.cpu().numpy(), oo).all(axis=1) & close(r_.cpu().numpy(), ro).all(axis=0)
            ok &= (d_.cpu().numpy() == do.astype(bool)).all(axis=0)
            assert (~ok).sum() <= 1
        for env in (a, b):
            ok = compare_engine_to_oracle(cfg, env.get_state(), eng.envs)
            assert (~ok).sum() <= 1
            if (~ok).any():
                env.set_state(oracle_state_to_engine(cfg, eng.envs))
    sa, sb = a.stats(), b.stats()
    assert abs(sa["episodes"] - sb["episodes"]) <= 2 and abs(sa["goals"] - sb["goals"]) <= 2
    assert abs(sa["episodes"] - eng.stats[2]) <= 2 and sa["episodes"] > 100
    a.close(); b.close()


def test_shard_invariance():
    """Envs keyed by global index: one handle of 1024 == two handles of 512."""
    full = sg.make("GoalContinuous3P-v0", 1024, seed=3, max_episode_steps=50)
    lo = sg.make("GoalContinuous3P-v0", 512, seed=3, first_env_index=0, max_episode_steps=50)
    hi = sg.make("GoalContinuous3P-v0", 512, seed=3, first_env_index=512, max_episode_steps=50)
    o = full.reset(); ol = lo.reset(); oh = hi.reset()
    assert torch.equal(o, torch.cat([ol, oh]))
    for _ in range(5):
        o, r, d, _ = full.step_many(K=20)
        ol, rl, dl, _ = lo.step_many(K=20)
        oh, rh, dh, _ = hi.step_many(K=20)
        assert torch.equal(o, torch.cat([ol, oh])) and torch.equal(r, torch.cat([rl, rh], dim=1))
        assert torch.equal(d, torch.cat([dl, dh], dim=1))
    s, sl, sh = full.stats(), lo.stats(), hi.stats()
    assert s["episodes"] == sl["episodes"] + sh["episodes"] and s["goals"] == sl["goals"] + sh["goals"]
    assert abs(s["return_sum"] - sl["return_sum"] - sh["return_sum"]) < 1e-6 * max(1, abs(s["return_sum"]))
    for e in (full, lo, hi):
        e.close()


@pytest.mark.parametrize("host_dtype", ["float32", "float64"])
def test_host_buffer_path_matches_device_path(host_dtype):
    """step(numpy) (packed host result block, one copy each way) == step(torch) on a twin handle: rewards and flags
    exact, observations exact (float64 block) or the float32 rounding of the device path's rows."""
    a = sg.make("GoalContinuous3P-v0", 1000, seed=11, host_obs_dtype=host_dtype)
    b = sg.make("GoalContinuous3P-v0", 1000, seed=11)
    a.reset(); b.reset()
    rng = np.random.RandomState(0)
    for _ in range(10):
        act = rng.uniform(-1, 1, size=(1000, 2)).astype(np.float32)
        oh, rh, dh, ih = a.step(act)
        od, rd, dd, idv = b.step(to_dev(act))
        want = od.cpu().numpy()
        assert oh.dtype == np.dtype(host_dtype)
        assert np.array_equal(oh, want.astype(host_dtype)) and np.array_equal(rh, rd.cpu().numpy())
        assert np.array_equal(dh, dd.cpu().numpy())
        assert np.array_equal(ih["TimeLimit.truncated"], idv["TimeLimit.truncated"].cpu().numpy())
    a.close(); b.close()


def test_host_path_is_ordered_after_the_callers_stream(oracle):
    """reset(); step(numpy) and mixed step(torch) / step(numpy) on one handle.  The host path runs on the
    handle's own non-blocking stream; it must start after the reset / device-path step enqueued on the caller's stream
    and the next device-path call must see its results.  Checked against the oracle twin transition by transition."""
    n, seed = 4096, 5
    env = sg.make("GoalContinuous3P-v0", n, seed=seed, host_obs_dtype="float64", max_episode_steps=40)
    cfg = oracle.config_from_id("GoalContinuous3P-v0", max_episode_steps=40)
    rng = np.random.RandomState(1)
    for rep in range(3):
        # env.seed(s); env.reset() replays what a fresh handle with seed s does (spaceship_env.py:92-94): the twin is new
        assert env.seed(seed + rep) == [seed + rep]
        twin = oracle.Engine(cfg, n, seed=seed + rep)
        od, oo = env.reset(), twin.reset()   # reset on the torch stream, host step right behind it
        assert close(od.cpu().numpy(), oo).all()
        for k in range(12):
            act = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
            oo, ro, do, _ = twin.step(act[None])
            if (k + rep) % 3 == 0:
                o, r, d, _ = env.step(to_dev(act))
                o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
            else:
                o, r, d, _ = env.step(act)
            ok = close(o, oo).all(axis=1) & close(r, ro[0]) & (d == do[0].astype(bool))
            assert (~ok).sum() == 0, f"rep {rep} step {k}: {(~ok).sum()} envs differ"
    # want_obs=False: reward | flags only, the observation stays on the device
    act = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    oo, ro, do, _ = twin.step(act[None])
    env.step_async(act, want_obs=False)
    o, r, d, _ = env.step_wait()
    assert o is None and close(r, ro[0]).all() and (d == do[0].astype(bool)).all()
    env.close()


def test_float32_observation_mode():
    a = sg.make("GoalContinuous4P-v0", 777, seed=2)
    b = sg.make("GoalContinuous4P-v0", 777, seed=2, obs_dtype="float32")
    oa, ob = a.reset(), b.reset()
    assert ob.dtype == torch.float32 and torch.equal(oa.float(), ob)
    for _ in range(5):
        oa, ra, da, _ = a.step_many(K=7)
        ob, rb, db, _ = b.step_many(K=7)
        assert torch.equal(oa.float(), ob) and torch.equal(ra, rb) and torch.equal(da, db)
    a.close(); b.close()


def test_vector_field_matches_oracle(oracle):
    n = 512
    for case in ("goal3p", "goal3p_accel", "kepler_hard_accel_rand"):
        cfg = oracle_cfg(oracle, case)
        env_id, kw = engine_kwargs(case)
        env = sg.make(env_id, n, seed=4, **kw)
        env.reset()
        st = env.get_state()
        act = np.random.RandomState(1).uniform(-1, 1, size=(n, 2)).astype(np.float32)
        vf = env.vector_field(to_dev(act)).cpu().numpy()
        L = oracle.lib()
        for i in range(0, n, 7):
            y = np.ascontiguousarray(st["ship"][i].cpu().numpy())
            pl = (st["planets"][i].cpu().numpy().ravel().copy() if cfg.task == 0 else np.zeros(4))
            a2 = np.array([(np.float32(act[i, 0]) + np.float32(1)) / np.float32(2), act[i, 1]], dtype=np.float64)
            out = np.zeros(6)
            L.so_rhs(C.byref(cfg), oracle._ptr(pl), oracle._ptr(a2), 1, oracle._ptr(y), oracle._ptr(out))
            assert close(vf[i], out).all()
        env.close()


@pytest.mark.parametrize("case", RESET_CASES)
def test_reset_distribution_matches_reference(case):
    """Reset states are drawn from another RNG than the reference's: compare distributions (4096 reference
    samples vs 65536 engine samples): means within 5 standard errors, spreads within 5 %, tile statistics."""
    ref = load("reset_samples", case)
    env_id, kw = engine_kwargs(case)
    n = 65536
    env = sg.make(env_id, n, seed=2024, **kw)
    env.reset()
    st = env.get_state()
    cols = {"ship": (st["ship"].cpu().numpy(), ref["state"].astype(np.float64))}
    if env.cfg.task == 0:
        P = env.cfg.n_planets
        cols["planets"] = (st["planets"].cpu().numpy().reshape(n, 2 * P), ref["planets"].reshape(-1, 2 * P).astype(np.float64))
        cols["goal"] = (st["goal"].cpu().numpy(), ref["goal"].astype(np.float64))
        if P > 1:  # tile statistics (the one-planet env has no tiling)
            tw = st["tile_word"].cpu().numpy()
            same = ((tw >> 2) & 63) == ((tw >> 8) & 63)
            assert abs(same.mean() - ref["same_tile"].mean()) < 0.03
        # objects never overlap and stay inside the world
        ship = st["ship"].cpu().numpy()[:, :2]
        pl = st["planets"].cpu().numpy()
        for p in range(P):
            d = np.linalg.norm(pl[:, p] - ship, axis=1)
            assert (d > env.cfg.planets_radius + env.cfg.ship_radius - 1e-12).all()
        assert (np.abs(pl) <= 1.5).all() and (np.abs(ship) <= 1.5).all()
    elif env.cfg.randomize:
        orb = st["orbit"].cpu().numpy()
        cols["orbit"] = (orb, np.stack([ref["orbit"][:, 1], ref["orbit"][:, 0]], axis=1).astype(np.float64))
    for name, (mine, theirs) in cols.items():
        for j in range(mine.shape[1]):
            a, b = mine[:, j], theirs[:, j]
            se = np.sqrt(a.var() / len(a) + b.var() / len(b)) + 1e-12
            assert abs(a.mean() - b.mean()) < 5 * se, (name, j, a.mean(), b.mean())
            assert abs(a.std() - b.std()) < 0.05 * b.std() + 1e-9, (name, j, a.std(), b.std())
            # two-sample Kolmogorov-Smirnov distance (the placements are multi-modal: quantiles are unstable)
            grid = np.sort(np.concatenate([a, b]))
            ks = np.max(np.abs(np.searchsorted(np.sort(a), grid, side="right") / len(a)
                               - np.searchsorted(np.sort(b), grid, side="right") / len(b)))
            assert ks < 2.2 * np.sqrt((len(a) + len(b)) / (len(a) * len(b))), (name, j, ks)
    env.close()


def test_full_size_rollout_properties(oracle):
    """BASELINE config 3 size: GoalContinuous3P-v0, 262144 envs, auto-reset. Size-independent checks."""
    n = 262144
    env = sg.make("GoalContinuous3P-v0", n, seed=7)
    obs = env.reset()
    assert obs.shape == (n, 15) and torch.isfinite(obs).all()
    total_done = 0
    for _ in range(4):
        obs, rew, done, info = env.step_many(K=50)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        total_done += int(done.sum())
    st = env.get_state()
    assert int(st["status"].abs().sum()) == 0
    ship = st["ship"]
    assert (ship[:, :2].abs() <= 1.5).all() and (ship[:, 2] >= 0).all() and (ship[:, 2] <= 2 * np.pi).all()
    assert (ship[:, 5].abs() <= 6).all()
    # unit heading vector and lidar geometry of the returned observations
    assert torch.allclose(obs[:, 2] ** 2 + obs[:, 3] ** 2, torch.ones(n, dtype=torch.float64, device=obs.device), atol=1e-12)
    goal_lidar = (st["goal"] - ship[:, :2]) * (2 / 3.0)
    same = ((obs[:, 13:15] - goal_lidar).abs() < 1e-12).all(dim=1)
    assert same.float().mean() > 0.99  # the rest reached their goal on this step: obs shows the old goal (quirk 3)
    s = env.stats()
    assert s["episodes"] == total_done
    assert abs(s["length_sum"] + float(st["elapsed"].sum()) - 200.0 * n) < 0.5  # every env step is accounted for
    frac = total_done / (200.0 * n)
    assert 0.01 < frac < 0.04  # ~2 % of random-policy steps end an episode
    # the benchmark configuration at its full size against the oracle: 1500 random transitions of the next step
    cfg = oracle.config_from_id("GoalContinuous3P-v0")
    assert _spot_check_against_oracle(oracle, "goal", env, cfg, 1500, np.random.RandomState(3)) == 0
    env.close()


@pytest.mark.parametrize("plan", ["SG_CONT1", "SG_FUSED", "SG_SUBBATCHES", "SG_STAGED"])
def test_alternative_step_plans_match_oracle(oracle, plan, monkeypatch):
    """The other launch plans of the split path (single-attempt continuation kernel, fused persistent
    hot+continuation kernel with in-kernel queue consumption, concurrent sub-batches replayed as a CUDA
    graph) stay in lockstep with the oracle twin as well."""
    monkeypatch.setenv(plan, "3" if plan == "SG_SUBBATCHES" else "1")
    n, T, seed = 3000, 60, 77
    cfg = oracle_cfg(oracle, "goal3p", max_episode_steps=25)
    eng = oracle.Engine(cfg, n, seed=seed)
    env = sg.make("GoalContinuous3P-v0", n, seed=seed, max_episode_steps=25, copy=False)
    assert close(env.reset().cpu().numpy(), eng.reset()).all()
    rng = np.random.RandomState(3)
    diverged = 0
    for t in range(T):
        a = random_actions(rng, 1, n, False)
        o_o, r_o, d_o, tr_o = eng.step(a)
        o_d, r_d, d_d, info = env.step(to_dev(a[0]))
        ok = close(o_d.cpu().numpy(), o_o).all(axis=1) & close(r_d.cpu().numpy(), r_o[0])
        ok &= (d_d.cpu().numpy() == d_o[0].astype(bool)) & compare_engine_to_oracle(cfg, env.get_state(), eng.envs)
        if (~ok).any():
            diverged += int((~ok).sum())
            env.set_state(oracle_state_to_engine(cfg, eng.envs))
    assert diverged <= 1
    env.close()


def _spot_check_against_oracle(oracle, case, env, cfg, n_check, rng):
    """Step the whole (large) env once and compare a random subset of transitions with the oracle."""
    n = env.num_envs
    st = env.get_state()
    idx = np.sort(rng.choice(n, size=n_check, replace=False))
    act = torch.rand((n, 2), device=env.device, generator=torch.Generator(device=env.device).manual_seed(5)) * 2 - 1
    act = act.float().contiguous()
    ship0 = st["ship"][idx].cpu().numpy()
    if cfg.task == 0:
        planets0 = st["planets"][idx].cpu().numpy()
        goal0 = st["goal"][idx].cpu().numpy()
        orbit0 = np.zeros((n_check, 3))
    else:
        planets0 = np.zeros((n_check, 2, 2))
        goal0 = np.zeros((n_check, 2))
        orb = st["orbit"][idx].cpu().numpy()
        orbit0 = np.stack([orb[:, 1], orb[:, 0], np.full(n_check, cfg.ref_orbit_a)], axis=1)  # (angle, ecc, a)
    elapsed0 = st["elapsed"][idx].cpu().numpy()
    obs, rew, done, info = env.step(act)
    a_np = act[idx].cpu().numpy()
    obs_np, rew_np, done_np = obs[idx].cpu().numpy(), rew[idx].cpu().numpy(), done[idx].cpu().numpy()
    bad = 0
    for k in range(n_check):
        r = oracle.step_transition(cfg, ship0[k], planets0[k], goal0[k], orbit0[k], a_np[k])
        dn = r["done"] or (elapsed0[k] + 1 >= cfg.max_episode_steps)
        ok = (dn == bool(done_np[k])) and close(rew_np[k], r["reward"]).all()
        if not dn:  # a done env returns its reset observation (other RNG draws than this one-shot oracle call)
            ok = ok and close(obs_np[k], r["obs"]).all()
        bad += 0 if ok else 1
    return bad


def test_config4_kepler_hard_randomize_accel_1m(oracle):
    """BASELINE config 4: KeplerEllipseHard-v0, randomize=True, angular-acceleration steering, 1M envs."""
    n = 1 << 20
    env = sg.make("KeplerEllipseHard-v0", n, seed=11, randomize=True, ship_steering=0)
    cfg = oracle.config_from_id("KeplerEllipseHard-v0", randomize=1, ship_steering=0)
    obs = env.reset()
    assert obs.shape == (n, 10) and torch.isfinite(obs).all()
    env.rollout_random(40)
    assert _spot_check_against_oracle(oracle, "kepler", env, cfg, 1500, np.random.RandomState(0)) == 0
    obs, rew, done, _ = env.step_many(K=20)
    st = env.get_state()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and int(st["status"].abs().sum()) == 0
    assert (rew > 0).all() and (rew <= 1).all()  # C / (penalties + C)
    orb = st["orbit"]
    assert (orb[:, 0] >= 0).all() and (orb[:, 0] < 0.7).all() and (orb[:, 1] >= 0).all() and (orb[:, 1] < 2 * np.pi).all()
    r = st["ship"][:, :2].norm(dim=1)
    assert (r >= 0.2 - 1e-9).all() and (r <= 3.0 + 1e-9).all() and (st["ship"][:, 5].abs() <= 6 + 1e-9).all()
    s = env.stats()
    assert s["episodes"] > 0 and abs(s["length_sum"] + float(st["elapsed"].sum()) - 61.0 * n) < 0.5
    env.close()


def test_config5_goal4p_4m_envs_sharded(oracle):
    """BASELINE config 5 (one shard of it): GoalContinuous4P-v0, 4M envs, as two shards keyed by global index."""
    n = 1 << 21
    cfg = oracle.config_from_id("GoalContinuous4P-v0")
    shards = [sg.make("GoalContinuous4P-v0", n, seed=5, first_env_index=k * n) for k in range(2)]
    bad = 0
    for k, env in enumerate(shards):
        obs = env.reset()
        assert obs.shape == (n, 17) and torch.isfinite(obs).all()
        env.rollout_random(30)
        bad += _spot_check_against_oracle(oracle, "goal", env, cfg, 1000, np.random.RandomState(k))
        st = env.get_state()
        assert int(st["status"].abs().sum()) == 0 and (st["ship"][:, :2].abs() <= 1.5).all()
    assert bad == 0
    # different global indices -> different streams (no duplicated envs across shards)
    a, b = shards[0].get_state()["ship"][:1000], shards[1].get_state()["ship"][:1000]
    assert not torch.equal(a, b)
    tot = sum(e.stats()["episodes"] for e in shards)
    assert tot > 0.01 * 31 * 2 * n * 0.5
    for e in shards:
        e.close()


def test_config5_goal4p_8m_envs_one_handle(oracle):
    """Top of BASELINE configs[4] on one GPU: ONE GoalContinuous4P-v0 handle of 8M envs (it splits
    itself into sub-batches): allocation, reset, 10 steps, integrity of every env, oracle spot check of one more step."""
    n = 1 << 23
    cfg = oracle.config_from_id("GoalContinuous4P-v0")
    env = sg.make("GoalContinuous4P-v0", n, seed=9, copy=False)
    obs = env.reset()
    assert obs.shape == (n, 17) and bool(torch.isfinite(obs).all())
    env.rollout_random(10)
    assert _spot_check_against_oracle(oracle, "goal", env, cfg, 500, np.random.RandomState(3)) == 0
    v = env.state_views()  # field-major views of the state block: no env-major copies of 8M rows
    assert int((v["status"] != 0).sum()) == 0, "sticky fault bits"
    assert bool(torch.isfinite(v["ship"]).all()) and bool((v["ship"][:2].abs() <= 1.5).all())
    assert int(v["elapsed"].max()) <= 11
    st = env.stats()
    assert 0 < st["episodes"] < 0.01 * n  # a fresh batch: few episodes end in its first 11 steps
    env.close()


@pytest.mark.parametrize("env_id", ["GoalContinuous3P-v0", "DoNotCrashContinuous-v0"])
def test_trajectory_dump_and_headless_replay(env_id, tmp_path):
    """recorded state -> .npz -> software-rendered frames (gym_space/rendering.py geometry)."""
    n, T = 64, 40
    env = sg.make(env_id, n, seed=11)
    rec = sg.TrajectoryRecorder(env, indices=[3, 17])
    env.reset()
    rec.start()
    rng = np.random.RandomState(0)
    for t in range(T):
        a = to_dev(rng.uniform(-1, 1, size=(n, 2)).astype(np.float32))
        obs, rew, done, _ = env.step(a)
        rec.record(a, rew, done)
    path = rec.save(str(tmp_path / "traj.npz"))
    d = np.load(path)
    assert d["ship"].shape == (T + 1, 2, 6) and d["action"].shape == (T, 2, 2) and d["done"].shape == (T, 2)
    assert close(d["ship"][-1], env.get_state()["ship"].cpu().numpy()[[3, 17]]).all()
    assert (d["action"][..., 0] >= 0).all() and (d["action"][..., 0] <= 1).all()  # engine in [0, 1]
    frames = sg.replay(path, env_slot=1, every=4)
    assert frames.shape == (11, 600, 600, 3) and frames.dtype == np.uint8
    # the ship disc sits where the recorded state says (white inside a dark outline)
    ws = float(d["world_size"])
    x, y = (d["ship"][-1, 1, :2] + ws / 2) * (600 / ws)  # T is a multiple of `every`: the last frame is the last state
    f = frames[-1]
    if 20 < x < 580 and 20 < y < 580:
        yy, xx = np.mgrid[-10:11, -10:11]
        disc = f[599 - int(y) + yy, int(x) + xx][(yy ** 2 + xx ** 2) <= 100]
        assert (disc == 255).all(axis=1).mean() > 0.6  # white body; the trace and the middle point are drawn over it
        ring = [f[599 - int(y + 15 * np.sin(k)), int(x + 15 * np.cos(k))].max() for k in np.linspace(2, 4.3, 12)]
        assert min(ring) < 170
    env.close()
