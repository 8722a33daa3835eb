"""ctypes binding of the C ABI in include/spacegym_b200.h (no torch types cross it)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPACEGYM_B200_LIB", os.path.join(HERE, "libspacegym_b200.so"))


class SgConfig(C.Structure):
    """Mirror of `sg_config`; field names are the reference's constructor kwargs."""
    _fields_ = [
        ("task", C.c_int32), ("n_planets", C.c_int32), ("ship_steering", C.c_int32), ("discrete", C.c_int32),
        ("max_episode_steps", C.c_int32), ("randomize", C.c_int32), ("auto_reset", C.c_int32),
        ("obs_dtype", C.c_int32), ("step_mode", C.c_int32), ("reserved0", C.c_int32),
        ("ship_mass", C.c_double), ("ship_moi", C.c_double), ("max_engine_force", C.c_double),
        ("max_thruster_force", C.c_double), ("world_size", C.c_double), ("max_abs_vel_angle", C.c_double),
        ("step_size", C.c_double),
        ("survival_reward_scale", C.c_double), ("goal_vel_reward_scale", C.c_double),
        ("safety_reward_scale", C.c_double), ("goal_sparse_reward", C.c_double), ("danger_zone", C.c_double),
        ("total_planets_mass", C.c_double),
        ("ref_orbit_a", C.c_double), ("ref_orbit_eccentricity", C.c_double), ("ref_orbit_angle", C.c_double),
        ("numerator_C", C.c_double), ("rad_penalty_C", C.c_double), ("act_penalty_C", C.c_double),
        ("kepler_planet_mass", C.c_double), ("kepler_planet_radius", C.c_double),
        ("kepler_border_radius", C.c_double),
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("hex_a", C.c_double), ("hex_height", C.c_double), ("hex_width", C.c_double), ("tiling_width", C.c_double),
        ("planets_radius", C.c_double), ("goal_radius", C.c_double), ("ship_radius", C.c_double),
    ]


class SgLayout(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_envs", "total_bytes", "ship", "planets", "goal", "orbit", "col_shift", "ep_return", "free_tiles",
        "tile_word", "elapsed", "rng_ctr", "status")]


class SgHostLayout(C.Structure):
    _fields_ = [("n_envs", C.c_int64), ("total_bytes", C.c_int64), ("obs", C.c_int64), ("reward", C.c_int64),
                ("flags", C.c_int64), ("obs_dtype", C.c_int32), ("obs_dim", C.c_int32)]


EXPORTS = [
    "sg_config_from_id", "sg_config_finalize", "sg_obs_dim", "sg_state_layout", "sg_create", "sg_destroy",
    "sg_last_error", "sg_get_layout", "sg_state_ptr", "sg_reset", "sg_step", "sg_reset_host", "sg_step_host", "sg_step_host_async", "sg_host_wait",
    "sg_host_block_layout", "sg_step_host_block_async", "sg_reseed", "sg_set_terminal_obs",
    "sg_vector_field", "sg_stats", "sg_stats_reset", "sg_step_count", "sg_set_step_count", "sg_kernel_launches",
    "sg_measure_fp64_peak", "sg_version", "sg_debug_kernel_times", "sg_debug_queue_check",
]

_lib = None


def lib():
    """Load the CUDA engine. There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m spacegym_b200.build` "
            "(spacegym_b200 has no CPU or PyTorch fallback path)")
    L = C.CDLL(LIB_PATH)
    vp, cp = C.c_void_p, C.POINTER(SgConfig)
    L.sg_config_from_id.argtypes = [C.c_char_p, cp]
    L.sg_config_finalize.argtypes = [cp]
    L.sg_obs_dim.argtypes = [cp]
    L.sg_state_layout.argtypes = [cp, C.c_int64, C.POINTER(SgLayout)]
    L.sg_create.argtypes = [cp, C.c_int64, C.c_int, C.c_uint64, C.c_int64, vp, C.POINTER(vp)]
    L.sg_destroy.argtypes = [vp]
    L.sg_destroy.restype = None
    L.sg_last_error.argtypes = [vp]
    L.sg_last_error.restype = C.c_char_p
    L.sg_get_layout.argtypes = [vp, C.POINTER(SgLayout)]
    L.sg_state_ptr.argtypes = [vp]
    L.sg_state_ptr.restype = vp
    L.sg_reset.argtypes = [vp, vp, vp, vp]
    L.sg_step.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.sg_reset_host.argtypes = [vp, vp]
    L.sg_step_host.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    L.sg_step_host_async.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    L.sg_host_wait.argtypes = [vp]
    L.sg_host_block_layout.argtypes = [vp, C.c_int, C.c_int, C.POINTER(SgHostLayout)]
    L.sg_step_host_block_async.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
    L.sg_reseed.argtypes = [vp, C.c_uint64]
    L.sg_set_terminal_obs.argtypes = [vp, vp]
    L.sg_debug_queue_check.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.sg_vector_field.argtypes = [vp, vp, vp, vp]
    L.sg_stats.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.sg_stats_reset.argtypes = [vp, vp]
    L.sg_step_count.argtypes = [vp]
    L.sg_step_count.restype = C.c_uint64
    L.sg_set_step_count.argtypes = [vp, C.c_uint64]
    L.sg_kernel_launches.argtypes = [vp]
    L.sg_kernel_launches.restype = C.c_int64
    L.sg_debug_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    L.sg_measure_fp64_peak.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double)]
    L.sg_version.restype = C.c_char_p
    _lib = L
    return L


class SgError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc != 0:
        msg = lib().sg_last_error(handle)
        raise SgError(f"spacegym_b200 error {rc}: {msg.decode() if msg else ''}")
