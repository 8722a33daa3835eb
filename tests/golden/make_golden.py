#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference; NumPy 2.3.5 / SciPy 1.18.1):

    python tests/golden/make_golden.py

It imports the reference package from /root/reference behind the gym shim in
oracle/gym/ and records, per env configuration,

  * transitions_<case>.npz : (state, planets, goal, orbit, raw action) -> (next state,
    observation, reward, done, goal-reached) harvested from random-action rollouts,
    over-sampling terminal events, goal hits and multi-step solves, with solver
    diagnostics (nfev, accepted steps, event index, end time) taken from solve_ivp;
  * resets_<case>.npz : GoalEnv/KeplerEnv `_reset` + successive `find_new_goal` calls
    driven by a tape RNG (explicit uniform/normal draws), which pins the draw ->
    placement mapping of hexagonal_tiling.py / goal.py / kepler.py exactly;
  * reset_samples_<case>.npz : resets with the reference's own RandomState, float32,
    for the statistical tests of the engine's reset distribution.

The fixtures travel to the GPU box; the reference does not.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import gym  # noqa: E402  (the shim)
import gym_space  # noqa: E402,F401
import gym_space.dynamic_model as dm  # noqa: E402
from gym_space.envs.goal import GoalDiscreteEnv  # noqa: E402
from gym_space.envs.kepler import KeplerDiscreteEnv  # noqa: E402

GOAL_DISCRETE_KW = dict(  # keyboard_agent.py:42-58
    n_planets=3, ship_steering=1, ship_moi=0.01, survival_reward_scale=0.2, goal_vel_reward_scale=5.0,
    safety_reward_scale=10.0, goal_sparse_reward=5.0, max_engine_force=1,
)
KEPLER_DISCRETE_KW = dict(  # keyboard_agent.py:10-27
    ship_steering=1, ship_moi=0.01, max_engine_force=0.4, reward_value=0, rad_penalty_C=2, numerator_C=0.01,
    act_penalty_C=0.5, step_size=0.07, randomize=False, ref_orbit_a=1.2, ref_orbit_eccentricity=0,
    ref_orbit_angle=0,
)

CASES = {
    "goal2p": ("GoalContinuous2P-v0", {}),
    "goal3p": ("GoalContinuous3P-v0", {}),
    "goal4p": ("GoalContinuous4P-v0", {}),
    "goal3p_accel": ("GoalContinuous3P-v0", {"ship_steering": 0}),
    "kepler_circle": ("KeplerCircleOrbit-v0", {}),
    "kepler_easy": ("KeplerEllipseEasy-v0", {}),
    "kepler_hard": ("KeplerEllipseHard-v0", {}),
    "kepler_hard_accel_rand": ("KeplerEllipseHard-v0", {"randomize": True, "ship_steering": 0}),
    "kepler_random": ("KeplerRandomOrbits-v0", {}),
    "goal3p_discrete": ("<GoalDiscrete>", GOAL_DISCRETE_KW),
    "kepler_discrete": ("<KeplerDiscrete>", KEPLER_DISCRETE_KW),
    # DoNotCrash cannot be constructed upstream: do_not_crash.py:19-21 calls ShipParams without its first field
    # `steering` (TypeError).  The fixtures supply exactly that one argument and nothing else (make_donotcrash).
    "donotcrash": ("<DoNotCrashContinuous>", {"ship_steering": 0}),
    "donotcrash_vel": ("<DoNotCrashContinuous>", {"ship_steering": 1}),
    "donotcrash_discrete": ("<DoNotCrashDiscrete>", {"ship_steering": 0}),
    # GoalEnv with n_planets=1 cannot run upstream (goal.py:134 puts the planet on the sampled goal position, goal.py:153
    # dereferences the missing tiling); the fixtures patch exactly those two lines (make_goal_one_planet)
    "goal1p": ("<GoalOnePlanet>", {}),
}


def make_donotcrash(discrete, ship_steering):
    import functools

    import gym_space.envs.do_not_crash as dnc
    from gym_space.ship_params import ShipParams, Steering
    saved = dnc.ShipParams
    dnc.ShipParams = functools.partial(ShipParams, Steering(ship_steering))  # the missing positional argument
    try:
        return (dnc.DoNotCrashDiscreteEnv if discrete else dnc.DoNotCrashContinuousEnv)()
    finally:
        dnc.ShipParams = saved


def make_goal_one_planet():
    """GoalContinuousEnv(n_planets=1) with the registry kwargs of GoalContinuous2P-v0 (gym_space/__init__.py:26-40) and the
    two statements that keep the one-planet configuration from running replaced; everything else is the reference's."""
    from gym_space.envs.goal import GoalContinuousEnv

    class GoalOnePlanet(GoalContinuousEnv):
        def _reset(self):  # goal.py:133-145; :134 `ship_pos, *planets_pos = ...` takes (ship, goal, planet) as (ship, *planets)
            ship_pos, _unused_goal_pos, planet_pos = self._sample_positions()
            self.planets[0].center_pos = planet_pos
            self._resample_goal()
            ship_angle = self._np_random.uniform(0, 2 * np.pi)
            velocities_xy = self._np_random.standard_normal(2) * 0.07
            max_abs_ang_vel = 0.7 * self.max_abs_vel_angle
            angular_velocity = self._np_random.standard_normal() * max_abs_ang_vel / 3
            angular_velocity = np.clip(angular_velocity, -max_abs_ang_vel, max_abs_ang_vel)
            self._ship_state.set(ship_pos, ship_angle, velocities_xy, angular_velocity)

        def _reward(self):  # goal.py:147-158; :153 reads self._hexagonal_tiling._debug (None here): threshold = goal_radius
            reward = (self.survival_reward_scale + self.goal_vel_reward_scale * self._goal_vel_reward2()
                      + self.safety_reward_scale * self._safety_reward_simple2())
            if np.linalg.norm(self.goal_pos - self._ship_state.pos_xy) < self.goal_radius:
                reward += self.goal_sparse_reward
                self._resample_goal()
            return reward

    return GoalOnePlanet(n_planets=1, ship_steering=1, ship_moi=0.01, survival_reward_scale=0.2, goal_vel_reward_scale=5.0,
                         safety_reward_scale=10.0, goal_sparse_reward=5.0, max_engine_force=0.4)


def make_env(case):
    env_id, kw = CASES[case]
    with contextlib.redirect_stdout(io.StringIO()):
        if env_id == "<GoalDiscrete>":
            env = GoalDiscreteEnv(**kw)
        elif env_id == "<KeplerDiscrete>":
            env = KeplerDiscreteEnv(**kw)
        elif env_id.startswith("<DoNotCrash"):
            env = make_donotcrash(env_id == "<DoNotCrashDiscrete>", **kw)
        elif env_id == "<GoalOnePlanet>":
            env = make_goal_one_planet()
        else:
            env = gym.make(env_id, **kw).unwrapped
    return env


# ---- solve_ivp spy: solver diagnostics of the last call -------------------------------
_last = {}
_orig_solve_ivp = dm.solve_ivp


def _spy(*a, **k):
    sol = _orig_solve_ivp(*a, **k)
    ev = -1
    for i, te in enumerate(sol.t_events):
        if len(te):
            ev = i
    _last.update(nfev=sol.nfev, n_steps=len(sol.t) - 1, event=ev, t_end=float(sol.t[-1]))
    return sol


dm.solve_ivp = _spy


def is_goal(env):
    return env.with_goal


def orbit_of(env):
    if is_goal(env) or not hasattr(env, "ref_orbit_angle"):
        return np.zeros(3)
    return np.array([env.ref_orbit_angle, env.ref_orbit_eccentricity, env.ref_orbit_a], dtype=np.float64)


def harvest_transitions(case, n_steps, n_keep, seed):
    env = make_env(case)
    env.seed(seed)
    np.random.seed(seed + 1)  # Kepler randomize uses the global RNG (kepler.py:257-259)
    act_rng = np.random.RandomState(seed + 2)
    discrete = not hasattr(env.action_space, "low")
    recs = []
    env.reset()
    t_ep = 0
    for _ in range(n_steps):
        if discrete:
            raw = int(act_rng.randint(6))
        else:
            raw = act_rng.uniform(-1, 1, size=2).astype(np.float32)
            k = act_rng.randint(12)  # a few boundary actions
            if k == 0:
                raw = np.round(raw).astype(np.float32)
            elif k == 1:
                raw[act_rng.randint(2)] = 0.0
        rec = dict(
            state0=env._ship_state._state_vec.copy(),
            planets=np.array([p.center_pos for p in env.planets], dtype=np.float64).copy(),
            goal=env.goal_pos.copy() if is_goal(env) else np.zeros(2),
            orbit=orbit_of(env),
            action=np.array([raw, 0], dtype=np.float32) if discrete else raw.copy(),
        )
        goal_before = env.goal_pos.copy() if is_goal(env) else None
        obs, reward, done, _ = env.step(raw)
        rec.update(
            state1=env._ship_state._state_vec.copy(),
            obs=np.asarray(obs, dtype=np.float64).copy(),
            reward=float(reward),
            done=bool(done),
            goal_reached=bool(is_goal(env) and not np.array_equal(goal_before, env.goal_pos)),
            nfev=_last["nfev"], n_steps=_last["n_steps"], event=_last["event"], t_end=_last["t_end"],
        )
        recs.append(rec)
        t_ep += 1
        if done or t_ep >= getattr(env, "max_episode_steps", 500):
            env.reset()
            t_ep = 0
    # category-balanced selection
    rng = np.random.RandomState(seed + 3)
    idx_done = [i for i, r in enumerate(recs) if r["done"]]
    idx_goal = [i for i, r in enumerate(recs) if r["goal_reached"]]
    idx_multi = [i for i, r in enumerate(recs) if r["n_steps"] >= 2 and not r["done"]]
    idx_rest = [i for i, r in enumerate(recs) if r["n_steps"] < 2 and not r["done"] and not r["goal_reached"]]
    keep = set()
    for pool, cap in ((idx_done, n_keep * 3 // 10), (idx_goal, n_keep // 8), (idx_multi, n_keep // 4)):
        pool = list(pool)
        rng.shuffle(pool)
        keep.update(pool[:cap])
    rest = list(idx_rest)
    rng.shuffle(rest)
    keep.update(rest[: max(0, n_keep - len(keep))])
    keep = sorted(keep)
    out = {k: np.array([recs[i][k] for i in keep]) for k in recs[0]}
    out["n_harvested"] = np.array(len(recs))
    return out


class TapeRNG:
    """Stand-in for the legacy RandomState driven by explicit draws (see module doc)."""

    def __init__(self, u, n):
        self.u, self.n, self.iu, self.i_n = u, n, 0, 0

    def _u(self):
        v = self.u[self.iu]
        self.iu += 1
        return v

    def uniform(self, low=0.0, high=1.0, size=None):
        if size is None:
            return low + (high - low) * self._u()
        return np.array([low + (high - low) * self._u() for _ in range(size)])

    def standard_normal(self, size=None):
        if size is None:
            v = self.n[self.i_n]
            self.i_n += 1
            return v
        v = np.array(self.n[self.i_n:self.i_n + size])
        self.i_n += size
        return v

    def randint(self, n):
        return min(int(self._u() * n), n - 1)

    def choice(self, n, size, replace=False):
        assert not replace
        arr = list(range(n))
        for i in range(size):
            j = min(i + int(self._u() * (n - i)), n - 1)
            arr[i], arr[j] = arr[j], arr[i]
        return np.array(arr[:size])


N_U, N_N, N_GOALS = 96, 4, 8


def harvest_resets(case, n, seed):
    env = make_env(case)
    rng = np.random.RandomState(seed)
    goal = is_goal(env)
    out = {k: [] for k in ("u", "n", "state", "planets", "goal", "orbit", "u_used", "tiles", "goals_seq",
                           "tiles_seq", "free_seq", "u_used_seq")}
    for _ in range(n):
        u = rng.uniform(size=N_U)
        nn = rng.standard_normal(size=N_N)
        tape = TapeRNG(u, nn)
        env._np_random = tape
        saved = np.random.uniform
        if goal:
            if env._hexagonal_tiling is not None:
                env._hexagonal_tiling._np_random = tape
        else:
            np.random.uniform = lambda: tape._u()  # kepler.py:258-259 (global RNG)
        try:
            env.reset()
        finally:
            np.random.uniform = saved
        out["u"].append(u)
        out["n"].append(nn)
        out["state"].append(env._ship_state._state_vec.copy())
        out["planets"].append(np.array([p.center_pos for p in env.planets], dtype=np.float64))
        out["goal"].append(env.goal_pos.copy() if goal else np.zeros(2))
        out["orbit"].append(orbit_of(env))
        out["u_used"].append(tape.iu)
        if goal and env._hexagonal_tiling is None:  # one planet: no tile state; successive goals around the planet
            out["tiles"].append(np.zeros(5 + 64, dtype=np.int64))
            gs, us = [], []
            for _g in range(N_GOALS):
                env._resample_goal()
                gs.append(env.goal_pos.copy())
                us.append(tape.iu)
            out["goals_seq"].append(gs)
            out["tiles_seq"].append(np.zeros((N_GOALS, 3), dtype=np.int64))
            out["free_seq"].append(np.zeros((N_GOALS, 64), dtype=np.int64))
            out["u_used_seq"].append(us)
        elif goal:
            ht = env._hexagonal_tiling
            free = list(ht._free_tiles_nrs) + [-1] * (64 - len(ht._free_tiles_nrs))
            out["tiles"].append(np.array([int(ht._case_b), int(ht._flip_xy), ht._ship_tile_nr, ht._goal_tile_nr,
                                          len(ht._free_tiles_nrs)] + free))
            gs, ts, fs, us = [], [], [], []
            for _g in range(N_GOALS):  # successive goals after the first (free-list multiset, quirk 8)
                env._resample_goal()
                gs.append(env.goal_pos.copy())
                ts.append([ht._ship_tile_nr, ht._goal_tile_nr, len(ht._free_tiles_nrs)])
                fs.append(list(ht._free_tiles_nrs) + [-1] * (64 - len(ht._free_tiles_nrs)))
                us.append(tape.iu)
            out["goals_seq"].append(gs)
            out["tiles_seq"].append(ts)
            out["free_seq"].append(fs)
            out["u_used_seq"].append(us)
        else:
            out["tiles"].append(np.zeros(5 + 64, dtype=np.int64))
            out["goals_seq"].append(np.zeros((N_GOALS, 2)))
            out["tiles_seq"].append(np.zeros((N_GOALS, 3), dtype=np.int64))
            out["free_seq"].append(np.zeros((N_GOALS, 64), dtype=np.int64))
            out["u_used_seq"].append(np.zeros(N_GOALS, dtype=np.int64))
    return {k: np.array(v) for k, v in out.items()}


def harvest_reset_samples(case, n, seed):
    env = make_env(case)
    env.seed(seed)
    np.random.seed(seed + 1)
    goal = is_goal(env)
    st, pl, gl, orb, same = [], [], [], [], []
    for _ in range(n):
        env.reset()
        st.append(env._ship_state._state_vec.copy())
        pl.append(np.array([p.center_pos for p in env.planets]))
        gl.append(env.goal_pos.copy() if goal else np.zeros(2))
        orb.append(orbit_of(env))
        ht = env._hexagonal_tiling if goal else None
        same.append(int(ht is not None and ht._goal_tile_nr == ht._ship_tile_nr))
    return dict(state=np.array(st, dtype=np.float32), planets=np.array(pl, dtype=np.float32),
                goal=np.array(gl, dtype=np.float32), orbit=np.array(orb, dtype=np.float32),
                same_tile=np.array(same, dtype=np.int8))


def main():
    n_steps = int(os.environ.get("GOLDEN_STEPS", 12000))
    n_keep = int(os.environ.get("GOLDEN_KEEP", 384))
    only = sys.argv[1:]
    for ci, case in enumerate(CASES):
        if only and case not in only:
            continue
        tr = harvest_transitions(case, n_steps, n_keep, seed=1000 + 17 * ci)
        np.savez_compressed(os.path.join(HERE, f"transitions_{case}.npz"), **tr)
        print(case, "transitions", len(tr["reward"]), "done", int(tr["done"].sum()), "goal",
              int(tr["goal_reached"].sum()), "multi", int((tr["n_steps"] >= 2).sum()), flush=True)
    for ci, case in enumerate(("goal2p", "goal3p", "goal4p", "kepler_circle", "kepler_random", "donotcrash", "goal1p")):
        if only and case not in only:
            continue
        rs = harvest_resets(case, 64, seed=5000 + ci)
        np.savez_compressed(os.path.join(HERE, f"resets_{case}.npz"), **rs)
        sm = harvest_reset_samples(case, 4096, seed=7000 + ci)
        np.savez_compressed(os.path.join(HERE, f"reset_samples_{case}.npz"), **sm)
        print(case, "resets", len(rs["state"]), "samples", len(sm["state"]), flush=True)
    # versions the fixtures were made with
    import scipy
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write(f"numpy {np.__version__}\nscipy {scipy.__version__}\npython {sys.version.split()[0]}\n")


if __name__ == "__main__":
    main()
