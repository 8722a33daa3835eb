/*
 * sg_oracle.c — CPU oracle (TEST INFRASTRUCTURE, see sg_oracle.h).
 *
 * Scalar, literal restatement of the reference hot path.  Each function cites the
 * reference (paths relative to the MIMUW-RL/space-gym tree) or SciPy 1.18.1
 * (`scipy:` = scipy/integrate/_ivp/) lines it follows.  Arithmetic is kept in the
 * reference's literal form (sqrt, /, atan2, cos, sin ...) on purpose: the CUDA
 * product path uses re-associated / trig-free forms, so agreement between the two
 * is a real check and not a tautology.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * Parity: pinned by tests/test_oracle_golden.py against tests/golden/ (npz fixtures)
 * (outputs of the reference run in the authoring container).
 */
#include "sg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_G 6.6743e-11 /* helpers.py:19 */
#define SO_EPS 2.220446049250313e-16
#define SO_TWO_PI (2 * 3.141592653589793)

/* ------------------------------------------------------------------------- */
/* configuration                                                              */
/* ------------------------------------------------------------------------- */

int so_obs_dim(const so_config* cfg) {
    /* spaceship_env.py:113-131, kepler.py:172-187 */
    if (cfg->task == SO_TASK_GOAL) return 7 + 2 * cfg->n_planets + 2;
    if (cfg->task == SO_TASK_DONOTCRASH) return 7; /* with_lidar=False, with_goal=False (do_not_crash.py:30-31) */
    return 10;
}

/* hexagonal_tiling.py:161-174 */
static void tiling_rows_cols_a(int min_tiles, double world_size, int* r_out, int* c_out, double* a_out) {
    double m = (double)min_tiles;
    double r_ = sqrt(72 * sqrt(3.0) * m - 6 * sqrt(3.0) + 12) / 12 - 1.0 / 4 + sqrt(3.0) / 12;
    int r = (int)ceil(r_);
    int c;
    for (;;) {
        c = (int)floor(2 * sqrt(3.0) * r / 3 - 1.0 / 3 + sqrt(3.0) / 3);
        if (r * c >= min_tiles) break;
        r += 1;
    }
    *r_out = r;
    *c_out = c;
    *a_out = 2 * sqrt(3.0) * world_size / (3 * (2 * r + 1));
}

void so_config_finalize(so_config* cfg) {
    if (cfg->task == SO_TASK_GOAL) {
        /* goal.py:10,43-47,62-72; hexagonal_tiling.py:15-48 */
        int P = cfg->n_planets;
        int n_objects = P + 2;
        int min_tiles = (P == 2) ? n_objects : (int)ceil(n_objects / 0.6);
        int r, c;
        double a;
        cfg->world_size = 3.0;
        cfg->step_size = 0.07;
        cfg->max_abs_vel_angle = 6.0;
        if (P == 1) { /* goal.py:79-82 (_init_one_planet): no tiling */
            cfg->rows = cfg->cols = 0;
            cfg->hex_a = cfg->hex_height = cfg->hex_width = cfg->tiling_width = 0;
            cfg->planet_radius[0] = 0.8;
            cfg->planet_mass[0] = 1e9;
            cfg->goal_radius = 0.2;
            cfg->ship_radius = 0.2;
            return;
        }
        tiling_rows_cols_a(min_tiles, cfg->world_size, &r, &c, &a);
        cfg->rows = r;
        cfg->cols = c;
        cfg->hex_a = a;
        cfg->hex_height = a * sqrt(3.0);
        cfg->hex_width = 2 * a;
        cfg->tiling_width = 3 * a * (c - 1) / 2 + 2 * a;
        {
            double pr = cfg->hex_height / 2;
            pr *= 0.75; /* PLANET_TILE_RATIO */
            for (int p = 0; p < P; ++p) {
                cfg->planet_radius[p] = pr;
                cfg->planet_mass[p] = 1e9 / P; /* goal.py:14,43 */
            }
            cfg->goal_radius = cfg->ship_radius = pr / 2;
        }
    } else if (cfg->task == SO_TASK_DONOTCRASH) {
        /* do_not_crash.py:10-32 */
        cfg->n_planets = 2;
        cfg->planet_mass[0] = 6e8;
        cfg->planet_radius[0] = 0.25;
        cfg->planet_mass[1] = 0.0;
        cfg->planet_radius[1] = 1.0;
        cfg->world_size = 2 * 1.0;
        cfg->step_size = 0.07;
        cfg->max_abs_vel_angle = 5.0;
        cfg->rows = cfg->cols = 0;
    } else {
        /* kepler.py:17-18,204-223 */
        cfg->n_planets = 2;
        cfg->planet_mass[0] = 6e8;
        cfg->planet_radius[0] = 0.2;
        cfg->planet_mass[1] = 0.0;
        cfg->planet_radius[1] = 3.0;
        cfg->world_size = 2 * 3.0;
        cfg->max_abs_vel_angle = 6;
        cfg->rows = cfg->cols = 0;
    }
}

int so_config_from_id(const char* id, so_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->ship_mass = 1;
    cfg->max_thruster_force = 0.05;
    cfg->max_episode_steps = 500;
    int goal_p = 0;
    if (!strcmp(id, "GoalContinuous2P-v0")) goal_p = 2;
    else if (!strcmp(id, "GoalContinuous3P-v0")) goal_p = 3;
    else if (!strcmp(id, "GoalContinuous4P-v0")) goal_p = 4;
    if (goal_p) {
        /* gym_space/__init__.py:26-72 */
        cfg->task = SO_TASK_GOAL;
        cfg->n_planets = goal_p;
        cfg->steering = 1;
        cfg->ship_moi = 0.01;
        cfg->survival_reward_scale = 0.2;
        cfg->goal_vel_reward_scale = 5.0;
        cfg->safety_reward_scale = 10.0;
        cfg->goal_sparse_reward = 5.0;
        cfg->max_engine_force = 0.4;
        cfg->danger_zone = 0.25;
        so_config_finalize(cfg);
        return 0;
    }
    if (!strcmp(id, "DoNotCrashContinuous-v0") || !strcmp(id, "DoNotCrashDiscrete-v0")) {
        /* gym_space/__init__.py:5-15, do_not_crash.py:19-21.  The reference omits ShipParams.steering
         * (TypeError at construction); the golden fixtures were made with that one argument supplied
         * (tests/golden/make_golden.py), default acceleration steering. */
        cfg->task = SO_TASK_DONOTCRASH;
        cfg->discrete = !strcmp(id, "DoNotCrashDiscrete-v0");
        cfg->steering = 0;
        cfg->ship_moi = 0.05;
        cfg->max_engine_force = 0.3;
        cfg->max_episode_steps = 300;
        so_config_finalize(cfg);
        return 0;
    }
    /* gym_space/__init__.py:76-146 */
    cfg->task = SO_TASK_KEPLER;
    cfg->steering = 1;
    cfg->ship_moi = 0.01;
    cfg->rad_penalty_C = 2;
    cfg->numerator_C = 0.01;
    cfg->act_penalty_C = 0.5;
    cfg->step_size = 0.07;
    cfg->max_engine_force = 0.4;
    cfg->ref_orbit_a = 1.2;
    if (!strcmp(id, "KeplerCircleOrbit-v0")) {
        cfg->ref_orbit_eccentricity = 0;
        cfg->ref_orbit_angle = 0;
    } else if (!strcmp(id, "KeplerEllipseEasy-v0")) {
        cfg->ref_orbit_eccentricity = 0.5;
        cfg->ref_orbit_angle = 0.8;
    } else if (!strcmp(id, "KeplerEllipseHard-v0")) {
        cfg->ref_orbit_eccentricity = 0.725;
        cfg->ref_orbit_angle = 3.925;
    } else if (!strcmp(id, "KeplerRandomOrbits-v0")) {
        cfg->randomize = 1;
        cfg->ref_orbit_eccentricity = 0.5; /* kepler.py:193-194 defaults */
        cfg->ref_orbit_angle = 3.75;
    } else {
        return -1;
    }
    so_config_finalize(cfg);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* physics: RHS and events                                                    */
/* ------------------------------------------------------------------------- */

/* dynamic_model.py:129-176 + helpers.py:22-35.  Mutates y[5] in velocity mode
 * exactly like the reference does. */
void so_rhs(const so_config* cfg, const double* planets, const double act[2], int is_f32,
            double y[6], double out[6]) {
    double engine = act[0], thruster = act[1];
    double fs, ft;
    if (is_f32) { /* NEP-50: float32 scalar * python float -> float32 (quirk 2) */
        fs = (double)((float)engine * (float)cfg->max_engine_force);
        ft = (double)((float)thruster * (float)cfg->max_thruster_force);
    } else {
        fs = engine * cfg->max_engine_force;
        ft = thruster * cfg->max_thruster_force;
    }
    double ang = y[2];
    double fx = -cos(ang) * fs; /* ship_external_force :168-176 */
    double fy = -sin(ang) * fs;
    for (int p = 0; p < cfg->n_planets; ++p) { /* helpers.gravity */
        double dx = planets[2 * p] - y[0];
        double dy = planets[2 * p + 1] - y[1];
        double d = sqrt(dx * dx + dy * dy);
        double ux = dx / d, uy = dy / d;
        double scalar = SO_G * cfg->ship_mass * cfg->planet_mass[p] / (d * d);
        fx += ux * scalar;
        fy += uy * scalar;
    }
    double ax = fx / cfg->ship_mass, ay = fy / cfg->ship_mass;
    double alpha = 0.0;
    if (cfg->steering == SO_STEER_ACCEL) alpha = ft / cfg->ship_moi;
    if (cfg->steering == SO_STEER_VELOCITY) {
        double w = is_f32 ? (double)((float)thruster * 5.0f) : thruster * 5.0;
        y[5] = w;
    }
    out[0] = y[3];
    out[1] = y[4];
    out[2] = y[5];
    out[3] = ax;
    out[4] = ay;
    out[5] = alpha;
}

/* dynamic_model.py:183-217; order: planets, world_max, world_min, angular velocity */
static int n_events(const so_config* cfg) { return cfg->n_planets + 3; }

static double event_value(const so_config* cfg, const double* planets, int i, const double y[6]) {
    int P = cfg->n_planets;
    if (i < P) {
        double dx = planets[2 * i] - y[0], dy = planets[2 * i + 1] - y[1];
        return sqrt(dx * dx + dy * dy) - cfg->planet_radius[i];
    }
    double half = cfg->world_size / 2;
    if (i == P) return fmin(half - y[0], half - y[1]);
    if (i == P + 1) return fmin(half + y[0], half + y[1]);
    return cfg->max_abs_vel_angle - fabs(y[5]);
}

/* ------------------------------------------------------------------------- */
/* SciPy RK45 (Dormand-Prince) with terminal events                           */
/* ------------------------------------------------------------------------- */

/* scipy:rk.py:538-565 */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
static const double RK_P[7][4] = {
    {1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

/* scipy:common.py:63-65 */
static double rms6(const double v[6]) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += v[i] * v[i];
    return sqrt(s) / sqrt(6.0);
}

typedef struct {
    const so_config* cfg;
    const double* planets;
    const double* act;
    int is_f32;
    int nfev;
} rhs_ctx;

static void fun(rhs_ctx* c, double y[6], double out[6]) {
    c->nfev += 1;
    so_rhs(c->cfg, c->planets, c->act, c->is_f32, y, out);
}

/* dense output, scipy:rk.py:178-180,715-737 */
typedef struct {
    double t_old, h, y_old[6], Q[6][4];
} dense_t;

static void dense_eval(const dense_t* d, double t, double y[6]) {
    double x = (t - d->t_old) / d->h;
    double p[4];
    p[0] = x;
    p[1] = p[0] * x;
    p[2] = p[1] * x;
    p[3] = p[2] * x;
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int j = 0; j < 4; ++j) s += d->Q[i][j] * p[j];
        y[i] = d->h * s + d->y_old[i];
    }
}

typedef struct {
    const so_config* cfg;
    const double* planets;
    const dense_t* dense;
    int event;
} root_ctx;

static double root_fun(root_ctx* r, double t) {
    double y[6];
    dense_eval(r->dense, t, y);
    return event_value(r->cfg, r->planets, r->event, y);
}

/* scipy.optimize.brentq as called from scipy:ivp.py:52-77 (xtol = rtol = 4 eps,
 * maxiter = 100): Brent's method in the bisection / secant / inverse-quadratic
 * form of scipy/optimize/Zeros/brentq.c (restated from the published algorithm).
 * Returns the root; *err != 0 on a sign error. */
static double brent_root(root_ctx* r, double xa, double xb, int* err) {
    const double xtol = 4 * SO_EPS, rtol = 4 * SO_EPS;
    double xpre = xa, xcur = xb, xblk = 0, fblk = 0, spre = 0, scur = 0;
    double fpre = root_fun(r, xpre);
    double fcur = root_fun(r, xcur);
    *err = 0;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) {
        *err = 1;
        return xb;
    }
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur;
            xcur = xblk;
            xblk = xpre;
            fpre = fcur;
            fcur = fblk;
            fblk = fpre;
        }
        double delta = (xtol + rtol * fabs(xcur)) / 2;
        double sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                double dpre = (fpre - fcur) / (xpre - xcur);
                double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = root_fun(r, xcur);
    }
    return xcur;
}

/* dynamic_model.py:94-125 -> scipy solve_ivp(RK45, t_span=(0, step_size), events)
 * scipy:ivp.py:578-760, scipy:rk.py:85-176, scipy:common.py:68-134, scipy:base.py:135-228 */
int so_make_step(const so_config* cfg, const double* planets, const double act[2], int is_f32,
                 double y[6], so_step_diag* diag, int32_t* status) {
    const double rtol = 1e-3, atol = 1e-6;
    const double T = cfg->step_size;
    const int NE = n_events(cfg);
    rhs_ctx ctx = {cfg, planets, act, is_f32, 0};
    double K[7][6];
    double f[6];
    int n_steps = 0, n_rej = 0, event_index = -1;
    /* smallest distance of a DISCRETE decision of this step from its threshold (test diagnostics: a divergence
     * between two correct implementations is only legitimate where this is at rounding level) */
    double margin = INFINITY;
#define SO_MARGIN(v) do { double m_ = fabs(v); if (m_ < margin) margin = m_; } while (0)

    /* RK45.__init__: f0 (mutates y[5] in velocity mode), then select_initial_step */
    fun(&ctx, y, f);
    double h_abs;
    {
        double scale[6], tmp[6], y1[6], f1[6];
        for (int i = 0; i < 6; ++i) scale[i] = atol + fabs(y[i]) * rtol;
        for (int i = 0; i < 6; ++i) tmp[i] = y[i] / scale[i];
        double d0 = rms6(tmp);
        for (int i = 0; i < 6; ++i) tmp[i] = f[i] / scale[i];
        double d1 = rms6(tmp);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        SO_MARGIN((d0 - 1e-5) / 1e-5);
        SO_MARGIN((d1 - 1e-5) / 1e-5);
        h0 = fmin(h0, T);
        for (int i = 0; i < 6; ++i) y1[i] = y[i] + h0 * 1.0 * f[i];
        fun(&ctx, y1, f1);
        for (int i = 0; i < 6; ++i) tmp[i] = (f1[i] - f[i]) / scale[i];
        double d2 = rms6(tmp) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5);
        h_abs = fmin(fmin(100 * h0, h1), T); /* max_step = inf */
    }

    double g[SO_MAX_PLANETS + 3], g_new[SO_MAX_PLANETS + 3];
    for (int i = 0; i < NE; ++i) g[i] = event_value(cfg, planets, i, y); /* scipy:ivp.py:651 */

    double t = 0.0;
    int done = 0;
    for (;;) { /* scipy:ivp.py:659 */
        /* ---- RungeKutta._step_impl, scipy:rk.py:111-176 ---- */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int rejected = 0;
        double y_new[6], f_new[6], h, t_new;
        for (;;) {
            if (h_abs < min_step) {
                *status |= SO_ERR_STEP_TOO_SMALL;
                goto finish; /* reference: assert ode_solution.success fails */
            }
            h = h_abs;
            t_new = t + h;
            if (t_new - T > 0) t_new = T;
            h = t_new - t;
            h_abs = fabs(h);
            /* rk_step, scipy:rk.py:14-71 */
            memcpy(K[0], f, sizeof(f));
            for (int s = 1; s < 6; ++s) {
                double ys[6];
                for (int i = 0; i < 6; ++i) {
                    double dot = 0;
                    for (int j = 0; j < s; ++j) dot += K[j][i] * RK_A[s][j];
                    ys[i] = y[i] + dot * h;
                }
                fun(&ctx, ys, K[s]);
            }
            for (int i = 0; i < 6; ++i) {
                double dot = 0;
                for (int j = 0; j < 6; ++j) dot += K[j][i] * RK_B[j];
                y_new[i] = y[i] + h * dot;
            }
            fun(&ctx, y_new, f_new);
            memcpy(K[6], f_new, sizeof(f_new));
            double err[6];
            for (int i = 0; i < 6; ++i) {
                double sc = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                double dot = 0;
                for (int j = 0; j < 7; ++j) dot += K[j][i] * RK_E[j];
                err[i] = dot * h / sc;
            }
            double error_norm = rms6(err);
            SO_MARGIN(error_norm - 1);
            if (error_norm < 1) {
                double factor;
                if (error_norm == 0) factor = 10;
                else factor = fmin(10, 0.9 * pow(error_norm, -0.2));
                if (rejected) factor = fmin(1, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, -0.2));
            rejected = 1;
            n_rej += 1;
        }
        n_steps += 1;
        double t_old = t;
        double y_old[6];
        memcpy(y_old, y, sizeof(y_old));
        t = t_new;
        memcpy(y, y_new, sizeof(y_new));
        memcpy(f, f_new, sizeof(f_new));
        int finished = (t - T >= 0); /* scipy:base.py:199 */

        /* ---- events, scipy:ivp.py:679-699 ---- */
        int any = 0;
        int active[SO_MAX_PLANETS + 3];
        for (int i = 0; i < NE; ++i) {
            g_new[i] = event_value(cfg, planets, i, y);
            SO_MARGIN(g_new[i]);
            int up = (g[i] <= 0) && (g_new[i] >= 0);
            int down = (g[i] >= 0) && (g_new[i] <= 0);
            active[i] = up || down;
            any |= active[i];
        }
        if (any) {
            dense_t d;
            d.t_old = t_old;
            d.h = t - t_old;
            memcpy(d.y_old, y_old, sizeof(y_old));
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 4; ++j) {
                    double s = 0;
                    for (int k = 0; k < 7; ++k) s += K[k][i] * RK_P[k][j];
                    d.Q[i][j] = s;
                }
            double best = INFINITY;
            for (int i = 0; i < NE; ++i) {
                if (!active[i]) continue;
                root_ctx rc = {cfg, planets, &d, i};
                int err = 0;
                double root = brent_root(&rc, t_old, t, &err);
                if (err) *status |= SO_ERR_BRENT_SIGN;
                if (root < best) { /* handle_events: earliest root wins */
                    best = root;
                    event_index = i;
                }
            }
            t = best;
            dense_eval(&d, t, y);
            done = 1;
            break;
        }
        memcpy(g, g_new, sizeof(double) * NE);
        if (finished) break;
    }
finish:
    y[2] = fmod(y[2], SO_TWO_PI); /* wrap_ship_angle :179-180, numpy floor-mod */
    if (y[2] != 0) {
        if (y[2] < 0) y[2] += SO_TWO_PI;
    } else {
        y[2] = 0.0;
    }
    if (diag) {
        diag->nfev = ctx.nfev;
        diag->n_steps = n_steps;
        diag->n_rejected = n_rej;
        diag->event_index = event_index;
        diag->t_end = t;
        diag->margin = margin;
    }
#undef SO_MARGIN
    for (int i = 0; i < 6; ++i)
        if (!isfinite(y[i])) *status |= SO_ERR_NONFINITE;
    return done;
}

/* ------------------------------------------------------------------------- */
/* observation                                                                */
/* ------------------------------------------------------------------------- */

/* spaceship_env.py:133-140 */
static void lidar(const double pos[2], const double obj[2], double radius, double world_size, double out[2]) {
    double vx = obj[0] - pos[0], vy = obj[1] - pos[1];
    double ang = atan2(vy, vx);
    ang = fmod(ang, SO_TWO_PI);
    if (ang != 0 && ang < 0) ang += SO_TWO_PI;
    double scale = (sqrt(vx * vx + vy * vy) - radius) * 2 / world_size;
    out[0] = cos(ang) * scale;
    out[1] = sin(ang) * scale;
}

/* spaceship_env.py:113-131, kepler.py:172-187 */
void so_make_obs(const so_config* cfg, const so_env* env, double* obs) {
    const double* y = env->y;
    obs[0] = y[0];
    obs[1] = y[1];
    obs[2] = cos(y[2]);
    obs[3] = sin(y[2]);
    obs[4] = y[3];
    obs[5] = y[4];
    obs[6] = y[5];
    if (cfg->task == SO_TASK_GOAL) {
        int P = cfg->n_planets;
        for (int p = 0; p < P; ++p)
            lidar(y, env->planets + 2 * p, cfg->planet_radius[p], cfg->world_size, obs + 7 + 2 * p);
        lidar(y, env->goal, 0.0, cfg->world_size, obs + 7 + 2 * P);
    } else if (cfg->task == SO_TASK_KEPLER) {
        obs[7] = env->orbit_angle;
        obs[8] = env->orbit_ecc;
        obs[9] = cfg->ref_orbit_a;
    }
}

/* ------------------------------------------------------------------------- */
/* rewards                                                                    */
/* ------------------------------------------------------------------------- */

static double norm2(double x, double y) { return sqrt(x * x + y * y); }

/* goal.py:147-165,204-227 (without the goal test, which the caller does) */
static double goal_reward_dense(const so_config* cfg, const so_env* env, const double last_xy[2]) {
    const double* y = env->y;
    double cur = norm2(env->goal[0] - y[0], env->goal[1] - y[1]);
    double last = norm2(env->goal[0] - last_xy[0], env->goal[1] - last_xy[1]);
    double gv = (last - cur) * 100;
    /* _safety_reward_simple2 */
    double sum_safety = 0;
    int closest = -1;
    double mindist = INFINITY;
    for (int p = 0; p < cfg->n_planets; ++p) {
        double dx = y[0] - env->planets[2 * p], dy = y[1] - env->planets[2 * p + 1];
        double dist = sqrt(dx * dx + dy * dy);
        if (dist < mindist) {
            closest = p;
            mindist = dist;
        }
    }
    if (closest >= 0) {
        double r = cfg->planet_radius[closest];
        double x0 = env->planets[2 * closest], y0 = env->planets[2 * closest + 1];
        if ((mindist - r) < cfg->danger_zone) {
            double px = last_xy[0] - x0, py = last_xy[1] - y0;
            double prev = sqrt(px * px + py * py);
            if (prev > mindist) sum_safety -= 100 * (prev - mindist);
        }
    }
    return cfg->survival_reward_scale + cfg->goal_vel_reward_scale * gv + cfg->safety_reward_scale * sum_safety;
}

/* kepler.py:43-156 */
static void rotate(const double v[2], double alpha, double out[2]) { /* :51-58 */
    double c = cos(alpha), s = sin(alpha);
    out[0] = c * v[0] + s * v[1];
    out[1] = -s * v[0] + c * v[1];
}

static double kepler_reward(const so_config* cfg, const so_env* env, const double act[2], int is_f32) {
    const double* y = env->y;
    double C = cfg->numerator_C;
    double a = cfg->ref_orbit_a, ecc = env->orbit_ecc, ang = env->orbit_angle;
    double b = sqrt(a * a * (1 - ecc * ecc));
    double c = sqrt(a * a - b * b);
    double pos[2] = {y[0], y[1]}, wz[2];
    /* _orbit_cur_rad :90-96 */
    rotate(pos, ang, wz);
    wz[0] = wz[0] - c;
    double cur_rad = norm2(wz[0], wz[1]);
    /* _orbit_target_vel :64-88 */
    double theta = atan2(wz[1], wz[0]);
    double ecos = ecc * cos(theta);
    double target_rad = b / sqrt(1 - ecos * ecos);
    double nwz = norm2(wz[0], wz[1]);
    double pw[2] = {wz[0] * target_rad / nwz, wz[1] * target_rad / nwz};
    double Vt[2];
    Vt[0] = -1 * a / b * pw[1];
    Vt[1] = 1 * b / a * pw[0];
    double r = norm2(pw[0] + c, pw[1] + 0);
    double GM = SO_G * cfg->planet_mass[0];
    double ov = sqrt(GM * (2 / r - 1 / a));
    double nvt = norm2(Vt[0], Vt[1]);
    Vt[0] = Vt[0] * ov / nvt;
    Vt[1] = Vt[1] * ov / nvt;
    double tv[2];
    rotate(Vt, -ang, tv);
    /* _dense_reward5 :111-150 */
    double rad_pen = fabs(cur_rad - target_rad);
    double vx_pen = fabs(tv[0] - y[3]);
    double vy_pen = fabs(tv[1] - y[4]);
    double act_term;
    if (is_f32) { /* float32 norm (no FMA) and float32 product, quirk 2 */
        float e = (float)act[0], t = (float)act[1];
        float ee = e * e, tt = t * t;
        float n = sqrtf(ee + tt);
        act_term = (double)((float)cfg->act_penalty_C * n);
    } else {
        act_term = cfg->act_penalty_C * sqrt(act[0] * act[0] + act[1] * act[1]);
    }
    return C / (cfg->rad_penalty_C * rad_pen + vx_pen + vy_pen + act_term + C);
}

/* ------------------------------------------------------------------------- */
/* RNG: Philox4x32-10 + draw streams                                          */
/* ------------------------------------------------------------------------- */

void so_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void so_draws_philox(so_draws* d, uint64_t seed, uint64_t env_index, uint32_t ctr) {
    memset(d, 0, sizeof(*d));
    d->seed = seed;
    d->env_index = env_index;
    d->ctr = ctr;
}

void so_draws_tape(so_draws* d, const double* u_tape, const double* n_tape) {
    memset(d, 0, sizeof(*d));
    d->u_tape = u_tape;
    d->n_tape = n_tape;
}

/* Engine stream 0: counter = (env_lo, env_hi, block, 0); a block gives 2 doubles. */
double so_next_uniform(so_draws* d) {
    if (d->u_tape) return d->u_tape[d->u_pos++];
    if (d->have_u) {
        d->have_u = 0;
        return d->ubuf;
    }
    uint32_t ctr[4] = {(uint32_t)d->env_index, (uint32_t)(d->env_index >> 32), d->ctr, 0u};
    uint32_t key[2] = {(uint32_t)d->seed, (uint32_t)(d->seed >> 32)};
    uint32_t o[4];
    so_philox4x32_10(ctr, key, o);
    d->ctr += 1;
    uint64_t a = ((uint64_t)o[1] << 32) | o[0];
    uint64_t b = ((uint64_t)o[3] << 32) | o[2];
    d->ubuf = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    d->have_u = 1;
    return (double)(a >> 11) * (1.0 / 9007199254740992.0);
}

/* Box-Muller on two engine uniforms; the reference uses numpy's legacy gauss, so
 * only the distribution (not the stream) is comparable. */
double so_next_normal(so_draws* d) {
    if (d->n_tape) return d->n_tape[d->n_pos++];
    if (d->have_n) {
        d->have_n = 0;
        return d->nbuf;
    }
    double u1 = 1.0 - so_next_uniform(d);
    double u2 = so_next_uniform(d);
    double r = sqrt(-2.0 * log(u1));
    double a = SO_TWO_PI * u2;
    d->nbuf = r * sin(a);
    d->have_n = 1;
    return r * cos(a);
}

/* Engine stream for synthetic actions: counter = (env_lo, env_hi, step_lo, 0x80000000|step_hi) */
void so_random_action(uint64_t seed, uint64_t env_index, uint64_t step, int discrete, float out[2]) {
    uint32_t ctr[4] = {(uint32_t)env_index, (uint32_t)(env_index >> 32), (uint32_t)step,
                       0x80000000u | (uint32_t)((step >> 32) & 0x7fffffffu)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    so_philox4x32_10(ctr, key, o);
    if (discrete) {
        out[0] = (float)(uint32_t)(((uint64_t)o[0] * 6u) >> 32);
        out[1] = 0.0f;
    } else {
        out[0] = (float)(o[0] >> 8) * (1.0f / 8388608.0f) - 1.0f;
        out[1] = (float)(o[1] >> 8) * (1.0f / 8388608.0f) - 1.0f;
    }
}

/* choice(n, size=k, replace=False): uniform ordered k-subset by partial Fisher-Yates
 * (numpy legacy draws permutation(n)[:k]: same distribution, different stream). */
static void choice_no_replace(so_draws* d, int n, int k, int* out) {
    int arr[64];
    for (int i = 0; i < n; ++i) arr[i] = i;
    for (int i = 0; i < k; ++i) {
        int j = i + (int)(so_next_uniform(d) * (n - i));
        if (j > n - 1) j = n - 1;
        int tmp = arr[i];
        arr[i] = arr[j];
        arr[j] = tmp;
        out[i] = arr[i];
    }
}

/* ------------------------------------------------------------------------- */
/* resets                                                                     */
/* ------------------------------------------------------------------------- */

/* hexagonal_tiling.py:136-158 */
static void tile_center(const so_config* cfg, const so_env* env, int tile, double out[2]) {
    int row = tile / cfg->cols, col = tile % cfg->cols;
    double zero_x = -cfg->world_size / 2 + cfg->hex_width / 2;
    double zero_y = cfg->world_size / 2 - cfg->hex_height / 2;
    if (env->case_b) zero_y -= cfg->hex_height / 2;
    double x_shift = col * 1.5 * cfg->hex_a + env->col_shift[col];
    double y_rows = -row * cfg->hex_height;
    double y_cols = -(col % 2) * cfg->hex_height / 2;
    if (env->case_b) y_cols *= -1;
    double cx = zero_x + x_shift, cy = zero_y + (y_rows + y_cols);
    if (env->flip_xy) {
        out[0] = cy;
        out[1] = cx;
    } else {
        out[0] = cx;
        out[1] = cy;
    }
}

/* helpers.py:38-46 (get_max_dist_in_direction): the positive candidates' minimum; IEEE division like NumPy's */
static double max_dist_in_direction(double max_pos, const double obj[2], const double unit[2]) {
    double cand[4] = {(max_pos - obj[0]) / unit[0], (-max_pos - obj[0]) / unit[0], (max_pos - obj[1]) / unit[1],
                      (-max_pos - obj[1]) / unit[1]};
    double best = INFINITY;
    int any = 0;
    for (int i = 0; i < 4; ++i)
        if (cand[i] > 0) {
            if (!any || cand[i] < best) best = cand[i];
            any = 1;
        }
    return best;
}

/* goal.py:90-99 (_sample_position_outside_one_planet) */
static void sample_outside_one_planet(const so_config* cfg, so_env* env, so_draws* d, const double planet[2],
                                      double clearance, double out[2]) {
    double max_obj_pos = cfg->world_size / 2 - clearance;
    double ang = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    double unit[2] = {cos(ang), sin(ang)};
    double dmax = max_dist_in_direction(max_obj_pos, planet, unit);
    double dmin = cfg->planet_radius[0] + clearance;
    if (!(dmin < dmax)) env->status |= SO_ERR_PLACEMENT; /* the reference asserts here */
    double dist = dmin + (dmax - dmin) * so_next_uniform(d);
    out[0] = planet[0] + unit[0] * dist;
    out[1] = planet[1] + unit[1] * dist;
}

/* hexagonal_tiling.py:95-128; one planet: goal.py:122-123 */
void so_find_new_goal(const so_config* cfg, so_env* env, so_draws* d) {
    if (cfg->n_planets == 1) {
        sample_outside_one_planet(cfg, env, d, env->planets, cfg->goal_radius, env->goal);
        return;
    }
    if (env->goal_tile >= 0) {
        if (env->n_free < SO_MAX_FREE) env->free_tiles[env->n_free++] = env->ship_tile;
        else env->status |= SO_ERR_FREE_OVERFLOW;
        env->ship_tile = env->goal_tile;
    }
    if (so_next_uniform(d) < 0.25) {
        env->goal_tile = env->ship_tile;
    } else {
        int k = env->n_free < 3 ? env->n_free : 3; /* MAX_GOAL_CANDIDATES */
        int cand[3];
        choice_no_replace(d, env->n_free, k, cand);
        int best_idx = -1, best = -1;
        int srow = env->ship_tile / cfg->cols, scol = env->ship_tile % cfg->cols;
        for (int i = 0; i < k; ++i) {
            int tnr = env->free_tiles[cand[i]];
            int taxi = abs(tnr / cfg->cols - srow) + abs(tnr % cfg->cols - scol);
            if (taxi > best) {
                best = taxi;
                best_idx = cand[i];
            }
        }
        env->goal_tile = env->free_tiles[best_idx];
        for (int i = best_idx; i + 1 < env->n_free; ++i) env->free_tiles[i] = env->free_tiles[i + 1];
        env->n_free -= 1;
    }
    /* _sample_disc_from_tile :130-134 + helpers.uniform_disk_distribution :48-53 */
    double cpos[2];
    tile_center(cfg, env, env->goal_tile, cpos);
    double R = cfg->hex_height / 2 - cfg->goal_radius;
    double ang = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    double rr = sqrt(so_next_uniform(d) * (R * R));
    env->goal[0] = cpos[0] + rr * cos(ang);
    env->goal[1] = cpos[1] + rr * sin(ang);
}

/* GoalEnv with n_planets = 1 (goal.py:79-82,90-123,133-145).  Upstream this configuration cannot run: `_reset`
 * unpacks (ship, goal, planet) as (ship, *planets) and so puts the planet on the sampled GOAL position, and
 * `_reward` dereferences `self._hexagonal_tiling._debug` on None.  Semantics chosen here ("the obvious
 * fix"), pinned by goldens generated from the reference with exactly these two lines patched
 * (tests/golden/make_golden.py, class GoalOnePlanet): the planet takes the sampled planet position (the first goal
 * sample is drawn and discarded: `_resample_goal` right after draws the goal the episode starts with), and the
 * goal-reached threshold is goal_radius. */
static void reset_goal_one_planet(const so_config* cfg, so_env* env, so_draws* d) {
    double max_pos = cfg->world_size / 2 - cfg->planet_radius[0];
    double clear = fmax(cfg->ship_radius, cfg->goal_radius);
    double pdist = 0 + ((max_pos - 2 * clear) - 0) * so_next_uniform(d);
    double pang = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    env->planets[0] = cos(pang) * pdist;
    env->planets[1] = sin(pang) * pdist;
    double ship[2], unused_goal[2];
    sample_outside_one_planet(cfg, env, d, env->planets, cfg->ship_radius, ship);
    sample_outside_one_planet(cfg, env, d, env->planets, cfg->goal_radius, unused_goal);
    env->y[0] = ship[0];
    env->y[1] = ship[1];
    env->goal_tile = -1;
    so_find_new_goal(cfg, env, d); /* goal.py:138 */
    env->y[2] = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    env->y[3] = so_next_normal(d) * 0.07;
    env->y[4] = so_next_normal(d) * 0.07;
    double maxw = 0.7 * cfg->max_abs_vel_angle;
    double w = so_next_normal(d) * maxw / 3;
    env->y[5] = fmin(fmax(w, -maxw), maxw);
}

static void reset_goal_env(const so_config* cfg, so_env* env, so_draws* d) {
    if (cfg->n_planets == 1) {
        reset_goal_one_planet(cfg, env, d);
        return;
    }
    /* HexagonalTiling.reset, hexagonal_tiling.py:53-93 */
    int P = cfg->n_planets, n_tiles = cfg->rows * cfg->cols;
    env->goal_tile = -1;
    env->case_b = so_next_uniform(d) < 0.5;
    env->flip_xy = so_next_uniform(d) < 0.5;
    double acc = 0;
    for (int c = 0; c < cfg->cols; ++c) {
        acc += so_next_uniform(d);
        env->col_shift[c] = acc;
    }
    double free_x = cfg->world_size - cfg->tiling_width;
    double sc = free_x / env->col_shift[cfg->cols - 1];
    for (int c = 0; c < cfg->cols; ++c) env->col_shift[c] *= sc;
    int tiles[SO_MAX_PLANETS + 1];
    int diag = 0;
    if (P == 2) diag = so_next_uniform(d) < 0.25;
    if (diag) {
        static const int cases[4][3] = {{1, 0, 3}, {2, 0, 3}, {0, 1, 2}, {3, 1, 2}};
        int k = (int)(so_next_uniform(d) * 4);
        if (k > 3) k = 3;
        for (int i = 0; i < 3; ++i) tiles[i] = cases[k][i];
    } else {
        choice_no_replace(d, n_tiles, P + 1, tiles);
    }
    env->ship_tile = tiles[0];
    env->n_free = 0;
    for (int i = 0; i < n_tiles; ++i) {
        int used = 0;
        for (int j = 0; j < P + 1; ++j) used |= (tiles[j] == i);
        if (!used) env->free_tiles[env->n_free++] = i;
    }
    double ang[SO_MAX_PLANETS + 1], pos[2];
    for (int i = 0; i < P + 1; ++i) ang[i] = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    for (int i = 0; i < P + 1; ++i) {
        double radius = (i == 0) ? cfg->ship_radius : cfg->planet_radius[i - 1];
        double R = cfg->hex_height / 2 - radius;
        double rr = sqrt(so_next_uniform(d) * (R * R));
        double cpos[2];
        tile_center(cfg, env, tiles[i], cpos);
        pos[0] = cpos[0] + rr * cos(ang[i]);
        pos[1] = cpos[1] + rr * sin(ang[i]);
        if (i == 0) {
            env->y[0] = pos[0];
            env->y[1] = pos[1];
        } else {
            env->planets[2 * (i - 1)] = pos[0];
            env->planets[2 * (i - 1) + 1] = pos[1];
        }
    }
    so_find_new_goal(cfg, env, d); /* goal.py:138 */
    /* goal.py:140-145 */
    env->y[2] = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    env->y[3] = so_next_normal(d) * 0.07;
    env->y[4] = so_next_normal(d) * 0.07;
    double maxw = 0.7 * cfg->max_abs_vel_angle;
    double w = so_next_normal(d) * maxw / 3;
    env->y[5] = fmin(fmax(w, -maxw), maxw);
}

static void reset_kepler_env(const so_config* cfg, so_env* env, so_draws* d) {
    /* kepler.py:233-267 */
    double lo = cfg->planet_radius[0] + 0.5, hi = cfg->planet_radius[1] - 0.5;
    double pang = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    double dist = lo + (hi - lo) * so_next_uniform(d);
    env->y[0] = cos(pang) * dist;
    env->y[1] = sin(pang) * dist;
    env->y[2] = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    if (cfg->randomize) {
        env->orbit_ecc = so_next_uniform(d) * 0.7;
        env->orbit_angle = so_next_uniform(d) * 2 * 3.141592653589793;
    }
    env->y[3] = so_next_normal(d) * 0.05;
    env->y[4] = so_next_normal(d) * 0.05;
    double maxw = 0.7 * cfg->max_abs_vel_angle;
    double w = so_next_normal(d) * maxw / 5;
    env->y[5] = fmin(fmax(w, -maxw), maxw);
    env->planets[0] = env->planets[1] = env->planets[2] = env->planets[3] = 0.0;
}

static void reset_donotcrash_env(const so_config* cfg, so_env* env, so_draws* d) {
    /* do_not_crash.py:34-45 */
    double pang = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    double lo = cfg->planet_radius[0] + 0.2, hi = cfg->planet_radius[1] - 0.15;
    double dist = lo + (hi - lo) * so_next_uniform(d);
    env->y[0] = cos(pang) * dist;
    env->y[1] = sin(pang) * dist;
    env->y[2] = 0 + (SO_TWO_PI - 0) * so_next_uniform(d);
    env->y[3] = so_next_normal(d) * 0.07;
    env->y[4] = so_next_normal(d) * 0.07;
    double maxw = 0.7 * cfg->max_abs_vel_angle;
    double w = so_next_normal(d) * maxw / 3;
    env->y[5] = fmin(fmax(w, -maxw), maxw);
    env->planets[0] = env->planets[1] = env->planets[2] = env->planets[3] = 0.0;
}

void so_reset(const so_config* cfg, so_env* env, so_draws* d) {
    if (cfg->task == SO_TASK_GOAL) reset_goal_env(cfg, env, d);
    else if (cfg->task == SO_TASK_DONOTCRASH) reset_donotcrash_env(cfg, env, d);
    else reset_kepler_env(cfg, env, d);
    env->elapsed = 0;
    env->ep_return = 0.0;
}

/* ------------------------------------------------------------------------- */
/* env.step                                                                   */
/* ------------------------------------------------------------------------- */

/* spaceship_env.py:68-78 (+189-202, 210-214) */
int so_step(const so_config* cfg, so_env* env, const float* raw_action, so_draws* draws,
            double* obs, double* reward, int32_t* goal_reached, so_step_diag* diag) {
    double act[2];
    int is_f32 = !cfg->discrete;
    if (cfg->discrete) {
        int a = (int)raw_action[0];
        static const double tab[6][2] = {{0, 0}, {1, 0}, {0, -1}, {0, 1}, {1, -1}, {1, 1}};
        act[0] = tab[a][0];
        act[1] = tab[a][1];
    } else {
        float e = (raw_action[0] + 1.0f) / 2.0f; /* float32 arithmetic */
        act[0] = (double)e;
        act[1] = (double)raw_action[1];
    }
    double last_xy[2] = {env->y[0], env->y[1]};
    int done = so_make_step(cfg, env->planets, act, is_f32, env->y, diag, &env->status);
    so_make_obs(cfg, env, obs); /* before _reward: old goal in obs (quirk 3) */
    int reached = 0;
    double r;
    if (cfg->task == SO_TASK_GOAL) {
        r = goal_reward_dense(cfg, env, last_xy);
        if (diag) { /* goal.py:154 is a discrete decision too */
            double m = fabs(norm2(env->goal[0] - env->y[0], env->goal[1] - env->y[1]) - cfg->goal_radius);
            if (m < diag->margin) diag->margin = m;
        }
        if (norm2(env->goal[0] - env->y[0], env->goal[1] - env->y[1]) < cfg->goal_radius) {
            r += cfg->goal_sparse_reward;
            reached = 1;
            if (draws) so_find_new_goal(cfg, env, draws);
        }
    } else if (cfg->task == SO_TASK_DONOTCRASH) {
        r = 100.0 / 300; /* do_not_crash.py:11,47-48: the class attribute max_episode_steps */
    } else {
        r = kepler_reward(cfg, env, act, is_f32);
    }
    *reward = r;
    if (goal_reached) *goal_reached = reached;
    return done;
}

/* ------------------------------------------------------------------------- */
/* batched engine twin (auto-reset + TimeLimit), layouts of include/spacegym_b200.h */
/* ------------------------------------------------------------------------- */

static int pick_threads(int n_threads) {
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    return n_threads;
}

void so_engine_reset(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                     int64_t first_env_index, double* obs, int n_threads) {
    int D = so_obs_dim(cfg);
    n_threads = pick_threads(n_threads);
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        so_env* e = &envs[i];
        so_draws d;
        memset(e, 0, sizeof(*e));
        e->orbit_ecc = cfg->ref_orbit_eccentricity;
        e->orbit_angle = cfg->ref_orbit_angle;
        so_draws_philox(&d, seed, (uint64_t)(first_env_index + i), 0);
        so_reset(cfg, e, &d);
        e->rng_ctr = d.ctr;
        if (obs) so_make_obs(cfg, e, obs + i * D);
    }
}

/* env.reset() on a LIVE batch (include/spacegym_b200.h: sg_reset with a mask; spaceship_env.py:59-66 called again
 * after an episode): the masked envs (mask == NULL: all) restart from the next Philox block of their own stream, as
 * the auto-reset inside so_engine_step does; the others keep their state and report their current observation. */
void so_engine_reset_live(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                          int64_t first_env_index, const uint8_t* mask, double* obs, int n_threads) {
    int D = so_obs_dim(cfg);
    n_threads = pick_threads(n_threads);
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        so_env* e = &envs[i];
        if (!mask || mask[i]) {
            so_draws d;
            so_draws_philox(&d, seed, (uint64_t)(first_env_index + i), e->rng_ctr);
            so_reset(cfg, e, &d);
            e->rng_ctr = d.ctr;
        }
        if (obs) so_make_obs(cfg, e, obs + i * D);
    }
}

void so_engine_step(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                    int64_t first_env_index, uint64_t step0, int K, const float* actions,
                    int auto_reset, double* obs, double* reward, uint8_t* done,
                    uint8_t* truncated, double* stats, int n_threads) {
    so_engine_step_term(cfg, envs, n, seed, first_env_index, step0, K, actions, auto_reset, obs, reward, done, truncated,
                        stats, NULL, n_threads);
}

/* so_engine_step that also reports, for every env that restarted in the LAST of the K steps, the observation of the
 * state its episode ended in (include/spacegym_b200.h: sg_set_terminal_obs); rows of other envs are not touched. */
void so_engine_step_term(const so_config* cfg, so_env* envs, int64_t n, uint64_t seed,
                         int64_t first_env_index, uint64_t step0, int K, const float* actions,
                         int auto_reset, double* obs, double* reward, uint8_t* done,
                         uint8_t* truncated, double* stats, double* term_obs, int n_threads) {
    int D = so_obs_dim(cfg);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    n_threads = pick_threads(n_threads);
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : s0, s1, s2, s3)
    for (int64_t i = 0; i < n; ++i) {
        so_env* e = &envs[i];
        uint64_t gidx = (uint64_t)(first_env_index + i);
        double o[SO_MAX_OBS];
        for (int k = 0; k < K; ++k) {
            float a[2];
            if (actions) {
                a[0] = actions[((int64_t)k * n + i) * 2];
                a[1] = actions[((int64_t)k * n + i) * 2 + 1];
            } else {
                so_random_action(seed, gidx, step0 + (uint64_t)k, cfg->discrete, a);
            }
            so_draws d;
            so_draws_philox(&d, seed, gidx, e->rng_ctr);
            double r;
            int32_t reached = 0;
            /* the goal resample of goal.py:155-157 is drawn below, not inside so_step: see the comment there */
            int dn = so_step(cfg, e, a, NULL, o, &r, &reached, NULL);
            e->elapsed += 1;
            int trunc = 0;
            if (e->elapsed >= cfg->max_episode_steps) { /* gym TimeLimit (0.21) */
                trunc = !dn;
                dn = 1;
            }
            /* Engine convention (the engine's streams are its own, SURVEY Appendix C: only the mapping from draws
             * to states is the reference's): an env that reaches its goal and restarts in the same step draws no
             * new goal -- the reset that follows overwrites goal and tiling -- so a reset always starts at the
             * env's current Philox block, whatever happened in the step. */
            if (reached && !(dn && auto_reset)) so_find_new_goal(cfg, e, &d);
            e->ep_return += r;
            s3 += reached;
            if (dn) {
                s0 += e->ep_return;
                s1 += e->elapsed;
                s2 += 1;
                if (auto_reset) { /* gym.vector SyncVectorEnv (0.21): obs := reset obs */
                    if (term_obs && k == K - 1) memcpy(term_obs + i * D, o, sizeof(double) * D);
                    d.have_u = d.have_n = 0; /* a reset starts on a fresh Philox block */
                    so_reset(cfg, e, &d);
                    so_make_obs(cfg, e, o);
                }
            }
            e->rng_ctr = d.ctr;
            if (reward) reward[(int64_t)k * n + i] = r;
            if (done) done[(int64_t)k * n + i] = (uint8_t)dn;
            if (truncated) truncated[(int64_t)k * n + i] = (uint8_t)trunc;
        }
        if (obs) memcpy(obs + i * D, o, sizeof(double) * D);
    }
    if (stats) {
        stats[0] += s0;
        stats[1] += s1;
        stats[2] += s2;
        stats[3] += s3;
    }
}
