"""Golden-case table shared by the oracle and the GPU parity tests (see tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# case -> (registered env id, reference-style kwargs overrides)
CASES = {
    "goal2p": ("GoalContinuous2P-v0", {}),
    "goal3p": ("GoalContinuous3P-v0", {}),
    "goal4p": ("GoalContinuous4P-v0", {}),
    "goal3p_accel": ("GoalContinuous3P-v0", {"ship_steering": 0}),
    "kepler_circle": ("KeplerCircleOrbit-v0", {}),
    "kepler_easy": ("KeplerEllipseEasy-v0", {}),
    "kepler_hard": ("KeplerEllipseHard-v0", {}),
    "kepler_hard_accel_rand": ("KeplerEllipseHard-v0", {"randomize": 1, "ship_steering": 0}),
    "kepler_random": ("KeplerRandomOrbits-v0", {}),
    # keyboard_agent.py:10-58 (discrete variants)
    "goal3p_discrete": ("GoalContinuous3P-v0", {"discrete": 1, "max_engine_force": 1.0}),
    "kepler_discrete": ("KeplerCircleOrbit-v0", {"discrete": 1}),
    # do_not_crash.py with the missing `steering` argument supplied (make_golden.py: make_donotcrash)
    "donotcrash": ("DoNotCrashContinuous-v0", {}),
    "donotcrash_vel": ("DoNotCrashContinuous-v0", {"ship_steering": 1}),
    "donotcrash_discrete": ("DoNotCrashDiscrete-v0", {}),
    # GoalEnv with one planet: goal.py:134,153 patched (make_golden.py: make_goal_one_planet)
    "goal1p": ("GoalContinuous2P-v0", {"n_planets": 1}),
}

RESET_CASES = ("goal2p", "goal3p", "goal4p", "kepler_circle", "kepler_random", "donotcrash", "goal1p")

# the parity bar: |delta| <= TOL * max(1, |ref|)  (BASELINE.json north_star: 1e-9 in FP64)
TOL = 1e-9


def load(kind, case):
    return np.load(os.path.join(GOLDEN, f"{kind}_{case}.npz"))


def close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))
