"""Env-ID registry: the ids and kwargs of gym_space/__init__.py:26-146 (+ the discrete
variants keyboard_agent.py:10-74 registers ad hoc), resolved to an `sg_config`.

`make_config(id, **kwargs)` accepts exactly the constructor kwargs of the reference classes
(GoalEnv.__init__ goal.py:18-31, KeplerEnv.__init__ kepler.py:189-203) with the same meaning,
and raises the same way Python would for an unknown one (TypeError).
"""
import ctypes as C

from ._lib import SgConfig, check, lib

GOAL_IDS = ("GoalContinuous2P-v0", "GoalContinuous3P-v0", "GoalContinuous4P-v0",
            "GoalDiscrete2-v0", "GoalDiscrete3-v0", "GoalDiscrete4-v0")
KEPLER_IDS = ("KeplerCircleOrbit-v0", "KeplerEllipseEasy-v0", "KeplerEllipseHard-v0", "KeplerRandomOrbits-v0",
              "KeplerDiscrete-v0")
# do_not_crash.py / gym_space/__init__.py:5-15.  Upstream the constructor raises TypeError (ShipParams without
# `steering`); here `ship_steering` is a keyword like in the other tasks (default 0 = acceleration).
DONOTCRASH_IDS = ("DoNotCrashContinuous-v0", "DoNotCrashDiscrete-v0")
ENV_IDS = GOAL_IDS + KEPLER_IDS + DONOTCRASH_IDS

# reference kwargs -> sg_config field (None: accepted and ignored, as in the reference)
_GOAL_KW = {
    "goal_vel_reward_scale": "goal_vel_reward_scale", "safety_reward_scale": "safety_reward_scale",
    "goal_sparse_reward": "goal_sparse_reward", "fixed_position": None, "danger_zone": "danger_zone",
    "survival_reward_scale": "survival_reward_scale", "n_planets": "n_planets", "ship_steering": "ship_steering",
    "ship_moi": "ship_moi", "max_engine_force": "max_engine_force", "renderer_kwargs": None,
}
_KEPLER_KW = {
    "randomize": "randomize", "ref_orbit_a": "ref_orbit_a", "ref_orbit_eccentricity": "ref_orbit_eccentricity",
    "ref_orbit_angle": "ref_orbit_angle", "reward_value": None, "numerator_C": "numerator_C",
    "rad_penalty_C": "rad_penalty_C", "act_penalty_C": "act_penalty_C", "step_size": "step_size",
    "ship_steering": "ship_steering", "ship_moi": "ship_moi", "max_engine_force": "max_engine_force",
}
_DONOTCRASH_KW = {"ship_steering": "ship_steering"}  # DoNotCrashEnv.__init__ takes no kwargs upstream
# engine-level options (not in the reference constructors)
_ENGINE_KW = {"max_episode_steps": "max_episode_steps", "auto_reset": "auto_reset", "discrete": "discrete",
              "obs_dtype": "obs_dtype", "step_mode": "step_mode"}


def make_config(env_id, **kwargs):
    cfg = SgConfig()
    rc = lib().sg_config_from_id(env_id.encode(), C.byref(cfg))
    if rc == -4:
        raise KeyError(f"No registered env with id: {env_id}")
    check(rc)
    table = dict((_GOAL_KW, _KEPLER_KW, _DONOTCRASH_KW)[cfg.task])
    table.update(_ENGINE_KW)
    for k, v in kwargs.items():
        if k not in table:
            raise TypeError(f"__init__() got an unexpected keyword argument '{k}'")
        field = table[k]
        if field is None:
            continue
        if k == "obs_dtype":
            v = {"float64": 0, "float32": 1, 0: 0, 1: 1}[str(v).replace("torch.", "") if not isinstance(v, int) else v]
        if k == "step_mode":
            v = {"auto": 0, "fused": 1, "split": 2, 0: 0, 1: 1, 2: 2}[v]
        if k == "ship_steering":
            v = int(getattr(v, "value", v))
            if v not in (0, 1):
                raise ValueError(f"{v} is not a supported Steering (0 = acceleration, 1 = velocity)")
        setattr(cfg, field, int(v) if isinstance(getattr(cfg, field), int) else float(v))
    if lib().sg_config_finalize(C.byref(cfg)) != 0:
        raise ValueError(f"unsupported configuration for {env_id}: {kwargs}")
    return cfg
