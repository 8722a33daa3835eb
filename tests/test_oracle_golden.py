"""CPU: pin the C oracle against the golden vectors produced by the reference itself."""
import ctypes as C

import numpy as np
import pytest

from golden_cases import CASES, RESET_CASES, close, load


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_transitions(oracle, case):
    env_id, kw = CASES[case]
    cfg = oracle.config_from_id(env_id, **kw)
    g = load("transitions", case)
    n = len(g["reward"])
    assert n >= 300 and g["done"].sum() >= 50
    worst = 0.0
    for i in range(n):
        r = oracle.step_transition(cfg, g["state0"][i], g["planets"][i], g["goal"][i], g["orbit"][i], g["action"][i])
        assert r["status"] == 0
        assert r["done"] == bool(g["done"][i]), (case, i)
        assert r["goal_reached"] == bool(g["goal_reached"][i]), (case, i)
        assert r["n_steps"] == g["n_steps"][i] and r["nfev"] == g["nfev"][i], (case, i)
        assert r["event"] == g["event"][i], (case, i)
        # the oracle is a literal restatement: hold it to 1e-12, far inside the 1e-9 bar
        for k in ("state1", "obs", "reward", "t_end"):
            assert close(r[k], g[k][i], 1e-12).all(), (case, i, k, r[k], g[k][i])
            worst = max(worst, float(np.max(np.abs(np.asarray(r[k]) - g[k][i]))))
    assert worst < 1e-12


@pytest.mark.parametrize("case", RESET_CASES)
def test_oracle_matches_reference_resets(oracle, case):
    """Same uniform/normal draws in -> same placement, tiles and goal sequence out."""
    env_id, kw = CASES[case]
    cfg = oracle.config_from_id(env_id, **kw)
    g = load("resets", case)
    L = oracle.lib()
    is_goal = cfg.task == 0
    P = cfg.n_planets
    for i in range(len(g["state"])):
        u = np.ascontiguousarray(g["u"][i])
        nn = np.ascontiguousarray(g["n"][i])
        d = oracle.Draws()
        L.so_draws_tape(C.byref(d), oracle._ptr(u), oracle._ptr(nn))
        env = oracle.Env()
        env.orbit_ecc, env.orbit_angle = cfg.ref_orbit_eccentricity, cfg.ref_orbit_angle
        L.so_reset(C.byref(cfg), C.byref(env), C.byref(d))
        assert d.u_pos == g["u_used"][i]
        assert close(np.array(env.y[:]), g["state"][i], 1e-13).all()
        if is_goal:
            assert close(np.array(env.planets[: 2 * P]), g["planets"][i].ravel(), 1e-13).all()
            assert close(np.array(env.goal[:]), g["goal"][i], 1e-13).all()
            tiled = P > 1  # the one-planet GoalEnv has no tiling (goal.py:79-82,90-123)
            t = g["tiles"][i]
            if tiled:
                assert (env.case_b, env.flip_xy, env.ship_tile, env.goal_tile, env.n_free) == tuple(t[:5])
                assert list(env.free_tiles[: env.n_free]) == list(t[5:5 + env.n_free])
            for k in range(g["goals_seq"].shape[1]):
                L.so_find_new_goal(C.byref(cfg), C.byref(env), C.byref(d))
                assert d.u_pos == g["u_used_seq"][i][k]
                if tiled:
                    assert (env.ship_tile, env.goal_tile, env.n_free) == tuple(g["tiles_seq"][i][k])
                    assert list(env.free_tiles[: env.n_free]) == list(g["free_seq"][i][k][: env.n_free])
                assert close(np.array(env.goal[:]), g["goals_seq"][i][k], 1e-13).all()
            assert env.status == 0
        else:
            assert close([env.orbit_angle, env.orbit_ecc, cfg.ref_orbit_a], g["orbit"][i], 1e-13).all()


def test_tiling_constants(oracle):
    """SURVEY.md §8a tiling table (values probed from the reference)."""
    want = {
        2: (2, 2, 0.6928203230275509, 1.2, 2.4248711305964283, 0.44999999999999996, 0.225),
        3: (3, 3, 0.4948716593053935, 0.8571428571428571, 2.4743582965269675, 0.3214285714285714, 0.1607142857142857),
        4: (4, 4, 0.3849001794597505, 0.6666666666666666, 2.501851166488378, 0.25, 0.125),
    }
    for P, (r, c, a, hh, tw, pr, gr) in want.items():
        cfg = oracle.config_from_id(f"GoalContinuous{P}P-v0")
        assert (cfg.rows, cfg.cols) == (r, c)
        assert cfg.hex_a == a and cfg.tiling_width == tw
        assert abs(cfg.hex_height - hh) < 1e-15
        assert abs(cfg.planet_radius[0] - pr) < 1e-15 and abs(cfg.goal_radius - gr) < 1e-15


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10 (Salmon et al., kat_vectors)."""
    L = oracle.lib()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.so_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_survey_known_answers(oracle):
    """SURVEY.md Appendix B rows A-C (reference outputs probed during the survey)."""
    cfg = oracle.config_from_id("GoalContinuous3P-v0")
    planets = [(0.9, 1.0), (-0.1, 0.7), (0.9, -0.6)]
    r = oracle.step_transition(cfg, [-1.0, 0.25, 3.5, -0.07, 0.07, 1.25], planets, (0.1, -0.9), (0, 0, 0), [0.3, 0.6])
    assert np.allclose(r["state1"], [-1.0042495241872404, 0.25518788279372, 3.71, -0.051699742311788116,
                                      0.07879242282841242, 3.0], rtol=0, atol=1e-14)
    assert abs(r["reward"] - -3.1432158120283438) < 1e-13 and not r["done"]
    r = oracle.step_transition(cfg, [-1.49, 0.25, 0.0, -0.5, 0.07, 0.0], planets, (0.1, -0.9), (0, 0, 0), [1.0, 0.0])
    assert r["done"] and abs(r["reward"] - -4.261535205109834) < 1e-13
    assert np.allclose(r["state1"], [-1.5, 0.2513900323115483, 0.0, -0.5076125245603876, 0.07006108881414727, 0.0],
                       rtol=0, atol=1e-14)
    cfg = oracle.config_from_id("KeplerEllipseHard-v0", ship_steering=0)
    r = oracle.step_transition(cfg, [-1.9, -0.6, 3.7, -0.1, 0.07, -0.7], np.zeros((2, 2)), (0, 0), (3.925, 0.725, 1.2),
                               [0.3, 1.0])
    assert np.allclose(r["state1"], [-1.9064314792303383, -0.5947628607702902, 3.6632500001825394,
                                      -0.08369974277110878, 0.0795363218657613, -0.3499999947845935], rtol=0, atol=1e-14)
    assert abs(r["reward"] - 0.00664986264606105) < 1e-15


def test_oracle_live_reset_continues_each_env_stream(oracle):
    """so_engine_reset_live (the twin of sg_reset on a running batch): masked envs restart from the next block of their
    own Philox stream — a second reset does not repeat the first placement — and unmasked envs keep their state."""
    cfg = oracle.config_from_id("GoalContinuous3P-v0")
    eng = oracle.Engine(cfg, 64, seed=2)
    o0 = eng.reset()
    ctr0 = eng.envs["rng_ctr"].copy()
    mask = np.zeros(64, dtype=np.uint8)
    mask[::3] = 1
    before = eng.envs.copy()
    o1 = eng.reset_live(mask)
    m = mask.astype(bool)
    assert (eng.envs["rng_ctr"][m] > ctr0[m]).all() and (eng.envs["rng_ctr"][~m] == ctr0[~m]).all()
    assert (o1[~m] == o0[~m]).all() and (np.abs(o1[m] - o0[m]).max(axis=1) > 0).all()
    assert (eng.envs["y"][~m] == before["y"][~m]).all() and (eng.envs["elapsed"][m] == 0).all()
    # a fresh engine reset restarts the streams: same placement as the very first reset
    assert (eng.reset() == o0).all()
