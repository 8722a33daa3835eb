/*
 * spacegym_b200.h — C ABI of the B200-native batched Space-Gym stepping engine.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (MIMUW-RL/space-gym) is pure Python behind the old gym.Env API and has no FFI of
 * its own, so each entry point cites the Python interface it replaces:
 *
 *   sg_config_from_id / sg_config_finalize   gym.make(id, **kwargs): gym_space/__init__.py:26-146,
 *                                            GoalEnv.__init__ goal.py:18-72, KeplerEnv.__init__
 *                                            kepler.py:189-231, HexagonalTiling.__init__
 *                                            hexagonal_tiling.py:15-48,161-174
 *   sg_create / sg_destroy                   SpaceshipEnv.__post_init__ spaceship_env.py:50-57 (+ env.seed :92-94)
 *   sg_reset / sg_reset_host                 SpaceshipEnv.reset spaceship_env.py:59-66 -> GoalEnv._reset goal.py:133-145
 *                                            / KeplerEnv._reset kepler.py:233-267
 *   sg_step / sg_step_host[_block_async]     SpaceshipEnv.step spaceship_env.py:68-78 -> ShipState.step
 *                                            dynamic_model.py:42-57,94-125 (scipy solve_ivp RK45 + events),
 *                                            _make_observation :113-131, GoalEnv._reward goal.py:147-158,
 *                                            KeplerEnv._reward kepler.py:111-156; gym TimeLimit + vector
 *                                            auto-reset (gym 0.21) around it
 *   sg_state_layout (+ caller-owned memory)  env._ship_state._state_vec, planet.center_pos, env.goal_pos,
 *                                            ref_orbit_*, HexagonalTiling fields hexagonal_tiling.py:42-43
 *   sg_vector_field                          SpaceshipEnv.vector_field spaceship_env.py:96-100
 *
 * Conventions: plain C, no torch types.  Pointers named d_* are DEVICE pointers on
 * the handle's device, h_* are HOST pointers.  `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  Device-pointer calls are asynchronous on
 * `stream`; *_host calls return after the results are in the host buffers.  Every
 * function returns 0 on success or a negative sg_status; sg_last_error() gives the
 * text.  A handle is bound to one device and is not thread-safe; distinct handles
 * are independent.  There is no CPU fallback: without a CUDA device sg_create fails.
 */
#ifndef SPACEGYM_B200_H
#define SPACEGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_MAX_PLANETS 4
#define SG_MAX_COLS 4
#define SG_FREE_WORDS 4 /* free-tile multiset: 64 x 4-bit entries */

/* SG_TASK_DONOTCRASH: gym_space/envs/do_not_crash.py (one planet inside a circular border, observation
 * [x, y, cos, sin, vx, vy, omega], reward 100/300 per step).  Upstream the class cannot be constructed
 * (ShipParams is called without `steering`, do_not_crash.py:19-21); here `ship_steering` is a parameter like
 * in the other tasks, default acceleration (the class sets a moment of inertia, which only that mode uses).
 * Its planet / border use the kepler_planet_* / kepler_border_radius fields. */
enum sg_task { SG_TASK_GOAL = 0, SG_TASK_KEPLER = 1, SG_TASK_DONOTCRASH = 2 };
enum sg_steering { SG_STEER_ACCELERATION = 0, SG_STEER_VELOCITY = 1 }; /* ship_params.py:5-8 */
enum sg_obs_dtype { SG_OBS_F64 = 0, SG_OBS_F32 = 1 };
/* How sg_step maps onto kernels.  SPLIT (= AUTO): every env step is a hot kernel + continuation and tail
 * kernels fed by device-side work queues (warps stay converged); K steps are 3 K launches.  FUSED: one kernel
 * runs all K substeps with the state in registers and the rare paths inline — one launch, but 3-4 x slower per
 * env step because every warp pays for every path. */
enum sg_step_mode { SG_STEP_AUTO = 0, SG_STEP_FUSED = 1, SG_STEP_SPLIT = 2 };

enum sg_status {
    SG_OK = 0,
    SG_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    SG_ERR_CUDA = -2,      /* CUDA runtime error (text in sg_last_error) */
    SG_ERR_NO_DEVICE = -3, /* no CUDA device: this library has no CPU path */
    SG_ERR_UNKNOWN_ID = -4
};

/* per-env sticky status bits (state field `status`) */
#define SG_ENV_STEP_TOO_SMALL 1u /* reference: assert ode_solution.success (dynamic_model.py:120) */
#define SG_ENV_BRENT_SIGN 2u     /* reference: brentq would raise ValueError */
#define SG_ENV_FREE_OVERFLOW 4u  /* free-tile multiset longer than 64 entries */
#define SG_ENV_NONFINITE 8u
#define SG_ENV_PLACEMENT 16u     /* one-planet GoalEnv: no room in the sampled direction (reference: assert, goal.py:96) */

/* Same names and meaning as the reference constructor kwargs. */
typedef struct sg_config {
    int32_t task;              /* sg_task */
    int32_t n_planets;         /* GoalEnv n_planets (2..4); Kepler: fixed planet + border */
    int32_t ship_steering;     /* sg_steering */
    int32_t discrete;          /* 1: Discrete(6) action path (spaceship_env.py:183-202) */
    int32_t max_episode_steps; /* gym registry TimeLimit; <= 0: none */
    int32_t randomize;         /* Kepler randomize */
    int32_t auto_reset;        /* 1: gym.vector semantics, done envs restart inside sg_step */
    int32_t obs_dtype;         /* sg_obs_dtype */
    int32_t step_mode;         /* sg_step_mode */
    int32_t reserved0;
    double ship_mass, ship_moi, max_engine_force, max_thruster_force;
    double world_size, max_abs_vel_angle, step_size;
    /* GoalEnv */
    double survival_reward_scale, goal_vel_reward_scale, safety_reward_scale;
    double goal_sparse_reward, danger_zone;
    double total_planets_mass;
    /* KeplerEnv */
    double ref_orbit_a, ref_orbit_eccentricity, ref_orbit_angle;
    double numerator_C, rad_penalty_C, act_penalty_C;
    double kepler_planet_mass, kepler_planet_radius, kepler_border_radius;
    /* derived by sg_config_finalize (hexagonal tiling geometry, radii) */
    int32_t rows, cols;
    double hex_a, hex_height, hex_width, tiling_width;
    double planets_radius, goal_radius, ship_radius;
} sg_config;

/* Byte offsets of the struct-of-arrays state inside the state block.  Every field is
 * an [rows][n_envs] array (env index fastest), 256-byte aligned. */
typedef struct sg_layout {
    int64_t n_envs;
    int64_t total_bytes;
    int64_t ship;       /* f64 [6][N]: x, y, theta, vx, vy, omega */
    int64_t planets;    /* f64 [2*P][N]: x0, y0, x1, y1, ... (Goal) */
    int64_t goal;       /* f64 [2][N] (Goal) */
    int64_t orbit;      /* f64 [2][N]: eccentricity, angle (Kepler) */
    int64_t col_shift;  /* f64 [cols][N] (Goal) */
    int64_t ep_return;  /* f64 [N] */
    int64_t free_tiles; /* u64 [4][N]: 4-bit tile numbers, entry k in word k/16 (Goal) */
    int64_t tile_word;  /* u32 [N]: bit0 case_b, bit1 flip_xy, 2..7 ship tile, 8..13 goal tile (63 = none), 14..20 n_free */
    int64_t elapsed;    /* u32 [N]: TimeLimit counter */
    int64_t rng_ctr;    /* u32 [N]: Philox block counter of the env's reset/goal stream */
    int64_t status;     /* u32 [N]: sticky SG_ENV_* bits */
} sg_layout;

/* Byte offsets inside the per-step HOST result block of sg_step_host_block_async: one contiguous block, so a step
 * is one device->host copy.  flags: bit 0 = done, bit 1 = truncated (TimeLimit; implies done). */
typedef struct sg_host_layout {
    int64_t n_envs;
    int64_t total_bytes;
    int64_t obs;     /* obs_dtype [N][D] (absent when want_obs = 0: obs == reward) */
    int64_t reward;  /* f64 [N] */
    int64_t flags;   /* u8 [N] */
    int32_t obs_dtype, obs_dim;
} sg_host_layout;

typedef struct sg_env sg_env; /* opaque handle */

/* ---- configuration ------------------------------------------------------------ */
int sg_config_from_id(const char* env_id, sg_config* cfg);
int sg_config_finalize(sg_config* cfg);
int sg_obs_dim(const sg_config* cfg);
int sg_state_layout(const sg_config* cfg, int64_t n_envs, sg_layout* out);

/* ---- lifetime ----------------------------------------------------------------- */
/* d_state: caller-owned device block of sg_state_layout().total_bytes (e.g. a torch
 * uint8 tensor), or NULL to let the library allocate it.  first_env_index is the
 * global index of env 0 of this shard: RNG streams are keyed by global index, so
 * results do not depend on how envs are split over GPUs. */
int sg_create(const sg_config* cfg, int64_t n_envs, int device, uint64_t seed,
              int64_t first_env_index, void* d_state, sg_env** out);
void sg_destroy(sg_env* env);
const char* sg_last_error(const sg_env* env); /* env may be NULL: last create error */
int sg_get_layout(const sg_env* env, sg_layout* out);
void* sg_state_ptr(const sg_env* env);

/* ---- hot path ----------------------------------------------------------------- */
/* Reset envs whose d_mask byte is non-zero (NULL: all).  d_obs: [N][D] or NULL. */
int sg_reset(sg_env* env, const uint8_t* d_mask, void* d_obs, void* stream);

/* K env steps (step_mode AUTO / SPLIT: 3 K kernel launches; FUSED: one launch).
 *   d_actions  f32 [K][N][2] (continuous: raw Box(-1,1) action; discrete: [..][0] = index 0..5),
 *              or NULL: synthetic uniform-random actions from the engine's counter RNG.
 *   d_obs      [N][D] observation after the last step (reset observation where the env
 *              restarted), obs_dtype; may be NULL.
 *   d_reward   f64 [K][N], d_done u8 [K][N], d_truncated u8 [K][N]; each may be NULL. */
int sg_step(sg_env* env, const float* d_actions, int K, void* d_obs, double* d_reward,
            uint8_t* d_done, uint8_t* d_truncated, void* stream);

/* Same calls with host buffers (pinned or pageable): actions go up, results come down. */
int sg_reset_host(sg_env* env, void* h_obs);
int sg_step_host(sg_env* env, const float* h_actions, int K, void* h_obs, double* h_reward,
                 uint8_t* h_done, uint8_t* h_truncated);
/* Asynchronous form (gym.vector step_async / step_wait): sg_step_host_async enqueues upload, step and
 * download on the handle's own stream and returns; sg_host_wait blocks until the host buffers hold the
 * results.  With pinned host buffers several handles overlap their copies and kernels. */
int sg_step_host_async(sg_env* env, const float* h_actions, int K, void* h_obs, double* h_reward,
                       uint8_t* h_done, uint8_t* h_truncated);
int sg_host_wait(sg_env* env);
/* Packed form of the same call (what BatchedSpaceEnv.step(numpy) uses): ONE host->device copy of the actions
 * f32 [N][2] (NULL: synthetic actions), one env step, ONE device->host copy of a contiguous result block laid out
 * by sg_host_block_layout.  obs_dtype is chosen per call (float32 — the dtype observation_space declares,
 * spaceship_env.py:110-111 — halves the block); want_obs = 0 leaves the observation on the device.  The work is
 * enqueued on the handle's own stream, ordered after everything already enqueued on `after_stream` (the stream the
 * caller uses for sg_reset / sg_step / state writes on this handle; NULL = legacy default stream).  Pinned host
 * memory keeps the copies asynchronous.  sg_host_wait() blocks until the block holds the results. */
int sg_host_block_layout(const sg_env* env, int obs_dtype, int want_obs, sg_host_layout* out);
int sg_step_host_block_async(sg_env* env, const float* h_actions, void* h_block, int obs_dtype, int want_obs,
                             void* after_stream);

/* RHS of the ship ODE for every env at its current state: d_out f64 [N][6]. */
int sg_vector_field(sg_env* env, const float* d_actions, double* d_out, void* stream);

/* ---- statistics and bookkeeping --------------------------------------------- */
/* out = {sum of finished-episode returns, sum of their lengths, finished episodes, goals reached}
 * since creation (or the last sg_stats_reset); synchronises `stream`. */
int sg_stats(sg_env* env, double out[4], void* stream);
int sg_stats_reset(sg_env* env, void* stream);
uint64_t sg_step_count(const sg_env* env);           /* env steps taken per env (drives the action RNG) */
int sg_set_step_count(sg_env* env, uint64_t steps);
/* env.seed(seed) (spaceship_env.py:92-94, goal.py:74-77): re-key every env's random streams and restart their
 * counters (and the step counter); the next reset / step draws from the new streams.  Synchronises the device. */
int sg_reseed(sg_env* env, uint64_t seed);
/* Terminal observations under auto-reset.  The reference's loop sees the last observation of an episode because the
 * caller resets (spaceship_env.py:68-78 returns it with done = True); with auto_reset the step returns the reset
 * observation instead (gym 0.21 SyncVectorEnv).  d_term_obs: [N][D] rows of the dtype of the step's d_obs (16-byte
 * aligned), or NULL to turn it off.  Every sg_step call that writes an observation (K = 1, or the last of K steps) then
 * also writes, for each env that restarted in that step, the observation of the state its episode ended in (terminal event or
 * TimeLimit); rows of other envs are not touched.  Split step path (the default) with device output buffers of the
 * handle's obs_dtype; the K-fused kernel and the host result block do not produce it.  Synchronises the device. */
int sg_set_terminal_obs(sg_env* env, void* d_term_obs);
int64_t sg_kernel_launches(const sg_env* env);       /* kernels launched by this handle */
/* Measurement aid, active when the handle was created with SG_TIMING=1 in the environment (sg_step then
 * synchronises inside every step): out[0..2] = mean milliseconds of the {hot, continuation, tail} kernels per
 * env step (CUDA events on the launching stream), out[3..5] = mean entries of the {continuation, event, post}
 * queues per step. */
int sg_debug_kernel_times(sg_env* env, double out[6], int reset);
/* Sanitizer substitute (compute-sanitizer is not available on the GPU pool): in a library built with
 * -DSG_CHECK_QUEUES every split step asserts on the device that the queue counters are zero when the step starts
 * and after its tail kernel, that every queue slot is below the capacity, that every queue entry is an env index
 * consumed once, and that every env left the step exactly once through {finish, terminal event}.  out[0..5] = number
 * of violations by kind (all zero = clean), out[6] = env steps checked, out[7] = 1.  In a normal build the call
 * returns SG_ERR_INVALID. */
int sg_debug_queue_check(sg_env* env, uint64_t out[8]);

/* ---- measurement helpers ----------------------------------------------------- */
/* Sustained FP64 FMA throughput of the device in TFLOP/s (2 flops per DFMA), measured
 * with a register-resident DFMA chain kernel for about `ms` milliseconds. */
int sg_measure_fp64_peak(int device, double ms, double* tflops_out);
const char* sg_version(void);

#ifdef __cplusplus
}
#endif
#endif
