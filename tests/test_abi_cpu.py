"""CPU: the C-ABI library loads, exports every symbol include/spacegym_b200.h declares, resolves the
registry like the reference, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import spacegym_b200 as sg
from spacegym_b200._lib import EXPORTS, LIB_PATH, SgConfig, SgLayout, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "spacegym_b200.h")).read()
    declared = set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS)
    L = C.CDLL(LIB_PATH)
    for name in sorted(declared):
        assert getattr(L, name) is not None
    assert b"sm_100a" in lib().sg_version()


def test_sg_config_struct_matches_header_size():
    # 10 int32 + 22 double + 2 int32 + 7 double
    assert C.sizeof(SgConfig) == 10 * 4 + 22 * 8 + 2 * 4 + 7 * 8
    assert C.sizeof(SgLayout) == 13 * 8


def test_registry_matches_reference_kwargs():
    """gym_space/__init__.py:26-146 and the SURVEY §8a tiling table."""
    want = {"GoalContinuous2P-v0": (2, 2, 0.6928203230275509, 0.44999999999999996, 13),
            "GoalContinuous3P-v0": (3, 3, 0.4948716593053935, 0.3214285714285714, 15),
            "GoalContinuous4P-v0": (4, 4, 0.3849001794597505, 0.25, 17)}
    for env_id, (r, c, a, pr, D) in want.items():
        cfg = sg.make_config(env_id)
        assert (cfg.rows, cfg.cols, cfg.hex_a) == (r, c, a)
        assert abs(cfg.planets_radius - pr) < 1e-15 and abs(cfg.goal_radius - pr / 2) < 1e-15
        assert lib().sg_obs_dim(C.byref(cfg)) == D
        assert (cfg.ship_steering, cfg.ship_moi, cfg.survival_reward_scale, cfg.goal_vel_reward_scale,
                cfg.safety_reward_scale, cfg.goal_sparse_reward, cfg.max_engine_force, cfg.max_episode_steps,
                cfg.step_size, cfg.world_size) == (1, 0.01, 0.2, 5.0, 10.0, 5.0, 0.4, 500, 0.07, 3.0)
    for env_id, (ecc, ang, rnd) in {"KeplerCircleOrbit-v0": (0, 0, 0), "KeplerEllipseEasy-v0": (0.5, 0.8, 0),
                                    "KeplerEllipseHard-v0": (0.725, 3.925, 0),
                                    "KeplerRandomOrbits-v0": (0.5, 3.75, 1)}.items():
        cfg = sg.make_config(env_id)
        assert (cfg.ref_orbit_eccentricity, cfg.ref_orbit_angle, cfg.randomize) == (ecc, ang, rnd)
        assert (cfg.ref_orbit_a, cfg.rad_penalty_C, cfg.numerator_C, cfg.act_penalty_C, cfg.step_size,
                cfg.world_size, cfg.max_episode_steps) == (1.2, 2, 0.01, 0.5, 0.07, 6.0, 500)
        assert lib().sg_obs_dim(C.byref(cfg)) == 10
    cfg = sg.make_config("KeplerEllipseHard-v0", randomize=True, ship_steering=0)  # BASELINE config 4
    assert cfg.randomize == 1 and cfg.ship_steering == 0
    for env_id, disc in (("DoNotCrashContinuous-v0", 0), ("DoNotCrashDiscrete-v0", 1)):  # do_not_crash.py:10-32
        cfg = sg.make_config(env_id)
        assert (cfg.task, cfg.discrete, cfg.ship_steering, cfg.ship_moi, cfg.max_engine_force, cfg.max_thruster_force,
                cfg.world_size, cfg.max_abs_vel_angle, cfg.step_size, cfg.max_episode_steps,
                cfg.kepler_planet_radius, cfg.kepler_border_radius) == (2, disc, 0, 0.05, 0.3, 0.05, 2.0, 5.0, 0.07,
                                                                        300, 0.25, 1.0)
        assert lib().sg_obs_dim(C.byref(cfg)) == 7
    assert sg.make_config("DoNotCrashContinuous-v0", ship_steering=1).ship_steering == 1
    with pytest.raises(TypeError):
        sg.make_config("DoNotCrashContinuous-v0", n_planets=2)


def test_registry_error_behaviour():
    with pytest.raises(KeyError):
        sg.make_config("NoSuchEnv-v0")
    with pytest.raises(TypeError):
        sg.make_config("GoalContinuous3P-v0", ref_orbit_a=1.0)  # a Kepler kwarg on a Goal env
    one = sg.make_config("GoalContinuous3P-v0", n_planets=1)  # goal.py:79-82: no tiling, radii 0.8 / 0.2 / 0.2
    assert (one.n_planets, one.rows, one.cols, one.planets_radius, one.goal_radius, one.ship_radius) == (1, 0, 0, 0.8, 0.2, 0.2)
    assert lib().sg_obs_dim(C.byref(one)) == 11
    with pytest.raises(ValueError):
        sg.make_config("GoalContinuous3P-v0", n_planets=0)
    with pytest.raises(ValueError):
        sg.make_config("GoalContinuous3P-v0", n_planets=5)
    with pytest.raises(ValueError):
        sg.make_config("GoalContinuous3P-v0", ship_steering=2)


def test_state_layout_is_aligned_and_disjoint():
    cfg = sg.make_config("GoalContinuous4P-v0")
    lay = SgLayout()
    assert lib().sg_state_layout(C.byref(cfg), 1000, C.byref(lay)) == 0
    offs = [lay.ship, lay.planets, lay.goal, lay.orbit, lay.col_shift, lay.ep_return, lay.free_tiles, lay.tile_word,
            lay.elapsed, lay.rng_ctr, lay.status, lay.total_bytes]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    assert lay.planets - lay.ship >= 6 * 8 * 1000 and lay.total_bytes >= (6 + 8 + 2 + 4 + 1 + 4) * 8000 + 16000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        sg.make("GoalContinuous3P-v0", 4)
    h = C.c_void_p()
    cfg = sg.make_config("GoalContinuous3P-v0")
    rc = lib().sg_create(C.byref(cfg), 4, 0, 0, 0, None, C.byref(h))
    assert rc == -3 and b"no CUDA device" in lib().sg_last_error(None)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "spacegym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|sg_oracle|so_engine|so_step", text), f


def test_bind_host_to_gpu_is_safe_without_a_gpu():
    """The NUMA helper never raises and never leaves the process with an empty CPU set."""
    before = os.sched_getaffinity(0)
    try:
        cpus = sg.bind_host_to_gpu(0)
        assert len(cpus) >= 1 and cpus <= before
    finally:
        os.sched_setaffinity(0, before)


def test_build_script_knows_every_part():
    """spacegym_b200/build.py compiles SG_PART=0..N_PARTS-1; the source must give every part an instantiation."""
    from spacegym_b200 import build
    src = open(os.path.join(ROOT, "spacegym_b200", "csrc", "sg_engine.cu")).read()
    for k in range(1, build.N_PARTS):
        assert re.search(rf"#elif SG_PART == {k}\ntemplate void launch_np<", src), k
    assert f"#elif SG_PART == {build.N_PARTS}" not in src


def test_spaces_sample_and_contain():
    """Box / Discrete stand-ins for the gym spaces the reference builds (spaceship_env.py:102-111,183-214)."""
    import numpy as np
    from spacegym_b200 import spaces
    box = spaces.Box(-np.ones(2, dtype=np.float32), np.ones(2, dtype=np.float32))
    rng = np.random.RandomState(0)
    for _ in range(50):
        x = box.sample(rng)
        assert x.dtype == np.float32 and box.contains(x)
    assert not box.contains(np.array([1.5, 0.0], dtype=np.float32)) and not box.contains(np.zeros(3, dtype=np.float32))
    hi = np.array([1.0, np.inf], dtype=np.float32)
    assert np.isfinite(spaces.Box(-hi, hi).sample(rng)).all()
    d = spaces.Discrete(6)
    assert all(d.contains(d.sample(rng)) for _ in range(50)) and not d.contains(6) and not d.contains(1.0)
