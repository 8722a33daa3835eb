/*
 * sg_oracle.h — CPU oracle for the Space-Gym step/reset path.
 *
 * TEST INFRASTRUCTURE.  This is a scalar, literal restatement in plain C of the
 * reference's algorithm (MIMUW-RL/space-gym + the SciPy 1.18.1 RK45/event/brentq
 * machinery it calls).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product package
 * (spacegym_b200/) never does.
 *
 * Parity status: PINNED against outputs of the reference itself run in the
 * authoring container (NumPy 2.3.5 / SciPy 1.18.1): tests/golden/ (npz fixtures), made by
 * tests/golden/make_golden.py.  The reference has no tests or golden vectors of
 * its own, so "the reference executed here" is the anchor.
 */
#ifndef SG_ORACLE_H
#define SG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_MAX_PLANETS 8
#define SO_MAX_COLS 8
#define SO_MAX_FREE 64
#define SO_MAX_OBS (9 + 2 * SO_MAX_PLANETS)

enum { SO_TASK_GOAL = 0, SO_TASK_KEPLER = 1, SO_TASK_DONOTCRASH = 2 }; /* goal.py, kepler.py, do_not_crash.py */
enum { SO_STEER_ACCEL = 0, SO_STEER_VELOCITY = 1 }; /* gym_space/ship_params.py:5-8 */

typedef struct so_config {
    int32_t task;
    int32_t n_planets;       /* bodies in `planets` (Goal: P; Kepler: 2 = planet + border) */
    int32_t steering;
    int32_t discrete;        /* 1: Discrete(6) actions, float64 action path */
    int32_t max_episode_steps;
    int32_t randomize;       /* Kepler: redraw ecc/angle on reset (kepler.py:257-259) */
    int32_t rows, cols;      /* hexagonal tiling (hexagonal_tiling.py:161-174) */
    double ship_mass, ship_moi, max_engine_force, max_thruster_force;
    double world_size, max_abs_vel_angle, step_size;
    double planet_mass[SO_MAX_PLANETS];
    double planet_radius[SO_MAX_PLANETS];
    /* GoalEnv (goal.py:18-72) */
    double survival_reward_scale, goal_vel_reward_scale, safety_reward_scale;
    double goal_sparse_reward, danger_zone, goal_radius, ship_radius;
    double hex_a, hex_height, hex_width, tiling_width;
    /* KeplerEnv (kepler.py:189-231) */
    double ref_orbit_a, ref_orbit_eccentricity, ref_orbit_angle;
    double numerator_C, rad_penalty_C, act_penalty_C;
} so_config;

/* Whole per-env state (what a checkpoint of the reference env would hold). */
typedef struct so_env {
    double y[6];                          /* x, y, theta, vx, vy, omega (dynamic_model.py:38-40) */
    double planets[2 * SO_MAX_PLANETS];   /* centre positions */
    double goal[2];
    double orbit_ecc, orbit_angle;        /* Kepler per-env reference orbit */
    /* hexagonal tiling episode state (hexagonal_tiling.py:42-43) */
    int32_t case_b, flip_xy;
    double col_shift[SO_MAX_COLS];
    int32_t ship_tile, goal_tile;         /* goal_tile < 0: None */
    int32_t n_free;
    int32_t free_tiles[SO_MAX_FREE];
    /* bookkeeping the batched engine adds */
    int32_t elapsed;                      /* TimeLimit counter */
    uint32_t rng_ctr;                     /* Philox block counter, stream 0 */
    double ep_return;
    int32_t status;                       /* sticky error bits */
} so_env;

typedef struct so_step_diag {
    int32_t nfev, n_steps, n_rejected, event_index; /* event_index = -1: none */
    double t_end;
    double margin; /* min |decision value - threshold| over the step's discrete decisions (accept/reject, event signs,
                      the h0 branch, goal reached); rounding-level margins are where two correct implementations may part */
} so_step_diag;

/* status bits */
#define SO_ERR_STEP_TOO_SMALL 1
#define SO_ERR_BRENT_SIGN 2
#define SO_ERR_FREE_OVERFLOW 4
#define SO_ERR_NONFINITE 8
#define SO_ERR_PLACEMENT 16 /* one-planet GoalEnv: no room in the sampled direction (the reference asserts, goal.py:96) */

/* draw source for resets: either a tape (pinning against the reference) or Philox */
typedef struct so_draws {
    const double* u_tape;   /* uniforms in [0,1) */
    const double* n_tape;   /* standard normals */
    int64_t u_pos, n_pos;
    /* Philox mode when u_tape == NULL */
    uint64_t seed, env_index;
    uint32_t ctr;
    double ubuf; int have_u;
    double nbuf; int have_n;
} so_draws;

int so_obs_dim(const so_config* cfg);

/* Fill cfg from a registered env id (gym_space/__init__.py:26-146) with the
 * reference's defaults; returns 0 or -1 (unknown id). Overrides are applied by
 * the caller on the struct, then so_config_finalize recomputes derived fields. */
int so_config_from_id(const char* env_id, so_config* cfg);
void so_config_finalize(so_config* cfg);

/* Physics: one env.step without reward/obs (dynamic_model.py:94-125). action is the
 * translated (engine, thruster) pair: is_f32 != 0 reproduces the float32 roundings. */
void so_rhs(const so_config* cfg, const double* planets, const double act[2], int is_f32,
            double y[6], double out[6]);
int so_make_step(const so_config* cfg, const double* planets, const double act[2], int is_f32,
                 double y[6], so_step_diag* diag, int32_t* status);

/* Full transition on an so_env, no auto reset: returns done; fills obs, reward,
 * goal_reached. When a goal is reached the new goal is drawn from `draws`
 * (may be NULL: goal left unchanged).  raw_action: 2 floats (continuous) or
 * raw_action[0] = discrete index. */
int so_step(const so_config* cfg, so_env* env, const float* raw_action, so_draws* draws,
            double* obs, double* reward, int32_t* goal_reached, so_step_diag* diag);

void so_make_obs(const so_config* cfg, const so_env* env, double* obs);
void so_reset(const so_config* cfg, so_env* env, so_draws* draws);
void so_find_new_goal(const so_config* cfg, so_env* env, so_draws* draws);

/* Counter-based RNG the batched engine uses (Philox4x32-10, Salmon et al. SC'11). */
void so_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void so_draws_philox(so_draws* d, uint64_t seed, uint64_t env_index, uint32_t ctr);
void so_draws_tape(so_draws* d, const double* u_tape, const double* n_tape);
double so_next_uniform(so_draws* d);
double so_next_normal(so_draws* d);
void so_random_action(uint64_t seed, uint64_t env_index, uint64_t step, int discrete, float out[2]);

/* Batched engine twin: n envs, K steps, auto-reset + TimeLimit, optional in-engine
 * random actions (actions == NULL).  Layouts follow include/spacegym_b200.h.
 * stats[4] += {sum return, sum length, episodes, goals}. n_threads <= 0: all cores. */
void so_engine_reset(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                     int64_t first_env_index, double* obs, int n_threads);
void so_engine_reset_live(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                          int64_t first_env_index, const uint8_t* mask, double* obs, int n_threads);
void so_engine_step(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                    int64_t first_env_index, uint64_t step0, int K, const float* actions,
                    int auto_reset, double* obs, double* reward, uint8_t* done,
                    uint8_t* truncated, double* stats, int n_threads);

void so_engine_step_term(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                         int64_t first_env_index, uint64_t step0, int K, const float* actions,
                         int auto_reset, double* obs, double* reward, uint8_t* done,
                         uint8_t* truncated, double* stats, double* term_obs, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
