"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libsg_oracle.so")

MAX_PLANETS, MAX_COLS, MAX_FREE = 8, 8, 64


class Config(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("n_planets", C.c_int32), ("steering", C.c_int32), ("discrete", C.c_int32),
        ("max_episode_steps", C.c_int32), ("randomize", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
        ("ship_mass", C.c_double), ("ship_moi", C.c_double), ("max_engine_force", C.c_double),
        ("max_thruster_force", C.c_double), ("world_size", C.c_double), ("max_abs_vel_angle", C.c_double),
        ("step_size", C.c_double),
        ("planet_mass", C.c_double * MAX_PLANETS), ("planet_radius", C.c_double * MAX_PLANETS),
        ("survival_reward_scale", C.c_double), ("goal_vel_reward_scale", C.c_double),
        ("safety_reward_scale", C.c_double), ("goal_sparse_reward", C.c_double), ("danger_zone", C.c_double),
        ("goal_radius", C.c_double), ("ship_radius", C.c_double),
        ("hex_a", C.c_double), ("hex_height", C.c_double), ("hex_width", C.c_double), ("tiling_width", C.c_double),
        ("ref_orbit_a", C.c_double), ("ref_orbit_eccentricity", C.c_double), ("ref_orbit_angle", C.c_double),
        ("numerator_C", C.c_double), ("rad_penalty_C", C.c_double), ("act_penalty_C", C.c_double),
    ]


class Env(C.Structure):
    _fields_ = [
        ("y", C.c_double * 6), ("planets", C.c_double * (2 * MAX_PLANETS)), ("goal", C.c_double * 2),
        ("orbit_ecc", C.c_double), ("orbit_angle", C.c_double),
        ("case_b", C.c_int32), ("flip_xy", C.c_int32), ("col_shift", C.c_double * MAX_COLS),
        ("ship_tile", C.c_int32), ("goal_tile", C.c_int32), ("n_free", C.c_int32),
        ("free_tiles", C.c_int32 * MAX_FREE),
        ("elapsed", C.c_int32), ("rng_ctr", C.c_uint32), ("ep_return", C.c_double), ("status", C.c_int32),
    ]


class Diag(C.Structure):
    _fields_ = [("nfev", C.c_int32), ("n_steps", C.c_int32), ("n_rejected", C.c_int32),
                ("event_index", C.c_int32), ("t_end", C.c_double), ("margin", C.c_double)]


class Draws(C.Structure):
    _fields_ = [("u_tape", C.POINTER(C.c_double)), ("n_tape", C.POINTER(C.c_double)),
                ("u_pos", C.c_int64), ("n_pos", C.c_int64), ("seed", C.c_uint64), ("env_index", C.c_uint64),
                ("ctr", C.c_uint32), ("ubuf", C.c_double), ("have_u", C.c_int), ("nbuf", C.c_double),
                ("have_n", C.c_int)]


ENV_DTYPE = np.dtype(Env)  # structured view for bulk access

_lib = None


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("sg_oracle.c", "sg_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.so_obs_dim.argtypes = [C.POINTER(Config)]
        L.so_config_from_id.argtypes = [C.c_char_p, C.POINTER(Config)]
        L.so_config_finalize.argtypes = [C.POINTER(Config)]
        L.so_rhs.argtypes = [C.POINTER(Config), dp, dp, C.c_int, dp, dp]
        L.so_make_step.argtypes = [C.POINTER(Config), dp, dp, C.c_int, dp, C.POINTER(Diag), C.POINTER(C.c_int32)]
        L.so_step.argtypes = [C.POINTER(Config), C.POINTER(Env), fp, C.POINTER(Draws), dp, dp,
                              C.POINTER(C.c_int32), C.POINTER(Diag)]
        L.so_make_obs.argtypes = [C.POINTER(Config), C.POINTER(Env), dp]
        L.so_reset.argtypes = [C.POINTER(Config), C.POINTER(Env), C.POINTER(Draws)]
        L.so_find_new_goal.argtypes = [C.POINTER(Config), C.POINTER(Env), C.POINTER(Draws)]
        L.so_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.so_draws_philox.argtypes = [C.POINTER(Draws), C.c_uint64, C.c_uint64, C.c_uint32]
        L.so_draws_tape.argtypes = [C.POINTER(Draws), dp, dp]
        L.so_next_uniform.argtypes = [C.POINTER(Draws)]
        L.so_next_uniform.restype = C.c_double
        L.so_next_normal.argtypes = [C.POINTER(Draws)]
        L.so_next_normal.restype = C.c_double
        L.so_random_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, fp]
        L.so_engine_reset.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p,
                                      C.c_int]
        L.so_engine_reset_live.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int]
        L.so_engine_step.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_uint64,
                                     C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int]
        L.so_engine_step_term.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_uint64,
                                          C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int]
        L.so_engine_step_term.restype = None
        for f in (L.so_rhs, L.so_config_finalize, L.so_make_obs, L.so_reset, L.so_find_new_goal,
                  L.so_philox4x32_10, L.so_draws_philox, L.so_draws_tape, L.so_random_action,
                  L.so_engine_reset, L.so_engine_reset_live, L.so_engine_step):
            f.restype = None
        _lib = L
    return _lib


def config_from_id(env_id, **overrides):
    """Registered id (gym_space/__init__.py) + reference-style kwargs overrides."""
    cfg = Config()
    if lib().so_config_from_id(env_id.encode(), C.byref(cfg)) != 0:
        raise KeyError(env_id)
    rename = {"ship_steering": "steering", "ref_orbit_eccentricity": "ref_orbit_eccentricity"}
    for k, v in overrides.items():
        k = rename.get(k, k)
        if not hasattr(cfg, k):
            raise TypeError(f"unknown override {k}")
        setattr(cfg, k, v)
    lib().so_config_finalize(C.byref(cfg))
    return cfg


def _ptr(a, typ=C.c_double):
    return a.ctypes.data_as(C.POINTER(typ))


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def step_transition(cfg, state, planets, goal, orbit, raw_action, goal_draws=None):
    """One env.step from an explicit (state, action): returns dict like the golden records."""
    L = lib()
    env = Env()
    for i in range(6):
        env.y[i] = state[i]
    pl = np.asarray(planets, dtype=np.float64).ravel()
    for i in range(len(pl)):
        env.planets[i] = pl[i]
    env.goal[0], env.goal[1] = goal[0], goal[1]
    env.orbit_angle, env.orbit_ecc = orbit[0], orbit[1]
    env.goal_tile = -1
    act = np.asarray(raw_action, dtype=np.float32).copy()
    obs = np.zeros(L.so_obs_dim(C.byref(cfg)))
    rew = C.c_double()
    reached = C.c_int32()
    diag = Diag()
    done = L.so_step(C.byref(cfg), C.byref(env), _ptr(act, C.c_float), goal_draws, _ptr(obs), C.byref(rew),
                     C.byref(reached), C.byref(diag))
    return dict(state1=np.array(env.y[:]), obs=obs, reward=rew.value, done=bool(done),
                goal_reached=bool(reached.value), nfev=diag.nfev, n_steps=diag.n_steps, event=diag.event_index,
                t_end=diag.t_end, margin=diag.margin, status=env.status)


class Engine:
    """CPU twin of the batched engine (same layouts, same Philox streams)."""

    def __init__(self, cfg, num_envs, seed=0, first_env_index=0, n_threads=0):
        self.cfg, self.n, self.seed, self.first = cfg, int(num_envs), int(seed), int(first_env_index)
        self.n_threads = n_threads
        self.D = lib().so_obs_dim(C.byref(cfg))
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        self.step_count = 0
        self.stats = np.zeros(4)

    def reset(self):
        obs = np.zeros((self.n, self.D))
        lib().so_engine_reset(C.byref(self.cfg), _vp(self.envs), self.n, self.seed, self.first, _vp(obs),
                              self.n_threads)
        self.step_count = 0
        return obs

    def reset_live(self, mask=None):
        """env.reset() of a batch that is already running (sg_reset): masked envs (None: all) restart from the next
        block of their own random stream; returns the observation of every env."""
        obs = np.zeros((self.n, self.D))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().so_engine_reset_live(C.byref(self.cfg), _vp(self.envs), self.n, self.seed, self.first, _vp(m), _vp(obs),
                                   self.n_threads)
        return obs

    def step(self, actions=None, K=1, auto_reset=True, want_obs=True, term_obs=None):
        """term_obs: float64 [n, D] array that receives the terminal observation of envs restarting in the last step."""
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.float32)
            if actions.ndim == 2:
                actions = actions[None]
            K = actions.shape[0]
        obs = np.zeros((self.n, self.D)) if want_obs else None
        rew = np.zeros((K, self.n))
        done = np.zeros((K, self.n), dtype=np.uint8)
        trunc = np.zeros((K, self.n), dtype=np.uint8)
        lib().so_engine_step_term(C.byref(self.cfg), _vp(self.envs), self.n, self.seed, self.first, self.step_count, K,
                                  _vp(actions), int(auto_reset), _vp(obs), _vp(rew), _vp(done), _vp(trunc),
                                  _vp(self.stats), _vp(term_obs), self.n_threads)
        self.step_count += K
        return obs, rew, done, trunc
