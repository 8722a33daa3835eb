// sg_device.cuh — device-side physics, integrator, rewards, observation and resets of the
// B200-native Space-Gym engine.  One CUDA thread owns one environment; everything here
// works on registers.  FP64 throughout, except the float32 roundings the reference
// inherits from NumPy-2 scalar promotion, which are reproduced with
// non-contractable __f*_rn intrinsics.
//
// Reference behaviour followed (MIMUW-RL/space-gym tree, SciPy 1.18.1 = `scipy:`):
//   RHS / forces          gym_space/dynamic_model.py:129-176, gym_space/helpers.py:22-35
//   events                gym_space/dynamic_model.py:183-217
//   RK45 + controller     scipy:rk.py:14-71,85-176,538-565; scipy:common.py:68-134
//   event loop, roots     scipy:ivp.py:52-157,651-699; scipy.optimize brentq (Brent 1973)
//   dense output          scipy:rk.py:178-180,715-737
//   observation           gym_space/envs/spaceship_env.py:113-140, kepler.py:172-187
//   rewards               gym_space/envs/goal.py:147-165,204-227; kepler.py:43-156
//   resets                gym_space/envs/goal.py:133-145, hexagonal_tiling.py:53-158, kepler.py:233-267
//
// The arithmetic is NOT a transcription: the ODE is integrated in Nystrom form (the force
// does not depend on velocity, the heading is a known polynomial in time), gravity uses
// rsqrt, lidars and the Kepler geometry are trig-free.  Control flow (accept/reject,
// step-size selection, event detection, earliest root) follows SciPy decision by decision.
#pragma once

#include <stdint.h>
#include <math.h>

namespace sg {

constexpr double kG = 6.6743e-11;  // helpers.py:19
constexpr double kPi = 3.141592653589793;
constexpr double kTwoPi = 2 * 3.141592653589793;
constexpr double kRtol = 1e-3, kAtol = 1e-6;  // solve_ivp defaults
constexpr double kEps = 2.220446049250313e-16;
constexpr double kInvSqrt6 = 0.40824829046386301637;

// ---- branch-free FP64 reciprocal / rsqrt / sqrt / divide --------------------------------------------
// The CUDA math library versions carry a slow-path CALL for denormal / infinite operands (argument
// marshalling, BSSY/BSYNC, register moves).  The operands here are ordinary normal numbers (distances,
// tolerances, norms), so the hardware approximation (MUFU.RCP64H / RSQ64H, ~2^-22) plus one cubically
// convergent correction gives <= 1-2 ulp without any branch.  None of these feed the float32 roundings
// of the action path (those keep the exact __f*_rn intrinsics).
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);  // y (1 + e + e^2)
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);  // 1 - x y^2
    return fma(y * e, fma(e, 0.375, 0.5), y);  // y (1 + e/2 + 3 e^2 / 8)
}
__device__ __forceinline__ double fast_sqrt(double x) {  // x > 0 (x == 0 handled by the caller where possible)
    const double y = fast_rsqrt(x);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);  // one Newton step on s: removes the rounding of x * y
}
__device__ __forceinline__ double fast_sqrt0(double x) { return x > 0.0 ? fast_sqrt(x) : 0.0; }
__device__ __forceinline__ double fast_div(double a, double b) {
    const double y = fast_rcp(b);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);  // residual correction
}

// min / max by one compare and a select.  CUDA's fmin / fmax implement the IEEE NaN rules and cost ~9 instructions on
// doubles here (DSETP with two predicates, quiet-bit LOP3, five moves / selects); the hot path calls them ~20 times per
// env step.  sel_min(a, b) / sel_max(a, b) return a unless b is strictly smaller / larger — Python's min / max, which
// is what SciPy's step-size rules use — so a NaN in b falls back to a: call sites pass the bound that must survive
// as a.  For ordered operands the results equal fmin / fmax.
#ifndef SG_FAST_MINMAX
#define SG_FAST_MINMAX 1
#endif
#if SG_FAST_MINMAX
__device__ __forceinline__ double sel_min(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double sel_max(double a, double b) { return b > a ? b : a; }
#else
__device__ __forceinline__ double sel_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double sel_max(double a, double b) { return fmax(a, b); }
#endif

// ---- sin / cos for a in [0, 2 pi] (wrapped headings, reset angles, Box-Muller angles) ------------------------
// a = q pi/2 + r with q = rint(a 2/pi) in 0..4, |r| <= pi/4 (two-constant Cody-Waite: q is tiny, so q * hi is
// exact); sin r through r^15, cos r through r^16: truncation errors < 5e-17 / 2e-18 at |r| = pi/4.
__device__ __forceinline__ void sincos_2pi(double a, double& sn, double& cs) {
    const double q = rint(a * 0.63661977236758134308);  // 2 / pi
    double r = fma(-q, 1.5707963267948965580, a);        // pi/2 (hi)
    r = fma(-q, 6.1232339957367660359e-17, r);           // pi/2 (lo)
    const double r2 = r * r;
    double ps = fma(r2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
    double pc = fma(r2, 1.0 / 20922789888000.0, -1.0 / 87178291200.0);
    ps = fma(ps, r2, -1.0 / 39916800.0);
    pc = fma(pc, r2, 1.0 / 479001600.0);
    ps = fma(ps, r2, 1.0 / 362880.0);
    pc = fma(pc, r2, -1.0 / 3628800.0);
    ps = fma(ps, r2, -1.0 / 5040.0);
    pc = fma(pc, r2, 1.0 / 40320.0);
    ps = fma(ps, r2, 1.0 / 120.0);
    pc = fma(pc, r2, -1.0 / 720.0);
    ps = fma(ps, r2, -1.0 / 6.0);
    pc = fma(pc, r2, 1.0 / 24.0);
    const double s = fma(ps * r2, r, r);
    pc = fma(pc, r2, -0.5);
    const double c = fma(pc, r2, 1.0);
    const int k = (int)q & 3;
    const double s1 = (k & 1) ? c : s, c1 = (k & 1) ? s : c;
    sn = (k & 2) ? -s1 : s1;
    cs = ((k + 1) & 2) ? -c1 : c1;
}


// out-of-line sincos for the cold paths (reset placement, terminal states): one copy of the range reduction
static __device__ __noinline__ void sincos_cold(double a, double* s, double* c) { sincos(a, s, c); }
// SG_FAST_SINCOS0 (option, off): quadrant-reduction sincos for the wrapped heading at the start of a step, library
// routine behind a call for headings injected from outside [0, 2 pi].  74 -> ~35 executed instructions, but the hot
// kernel went from 36.4 to 38.6 us (5.89e9 vs 6.05e9 env-steps/s): the call in the kernel body costs more than the
// shorter sequence returns.  Select-based min / max alone: 6.08e9 vs 6.05e9, adopted.
#ifndef SG_FAST_SINCOS0
#define SG_FAST_SINCOS0 0
#endif

// ---- x^(1/5) and x^(-1/10) without pow ----------------------------------------------------------------
// The step-size rules need (0.01 / d)^(1/5) (scipy:common.py:131) and err^(-1/5) = (err^2)^(-1/10)
// (scipy:rk.py:150-161).  CUDA's pow is a ~200-instruction serial chain behind a CALL that a warp walks
// whenever any of its lanes needs the value.  Here: float seed from MUFU.LG2/EX2 (relative error ~3e-7)
// and two division-free Newton steps on the inverse root (quadratic: 3e-7 -> ~5e-13 -> < 1 ulp), i.e.
// ~20 FP64 instructions with no branch.  Results agree with a correctly rounded pow to ~2 ulp.
#ifndef SG_LIBM_POW
__device__ __forceinline__ double fifth_root(double x) {  // x in [1e-30, 1e30]
    double z = (double)exp2f(-0.2f * __log2f((float)x));  // z ~ x^(-1/5)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double z2 = z * z, z4 = z2 * z2;
        const double r = fma(-(x * z), z4, 1.0);  // 1 - x z^5
        z = fma(z * 0.2, r, z);                   // Newton on z^-5 = x
    }
    const double z2 = z * z;
    return x * (z2 * z2);  // x^(1/5) = x z^4
}
__device__ __forceinline__ double inv_tenth_root(double x) {  // x^(-1/10), x in [1e-30, 1e30]
    double z = (double)exp2f(-0.1f * __log2f((float)x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
        const double r = fma(-(x * z2), z8, 1.0);  // 1 - x z^10
        z = fma(z * 0.1, r, z);
    }
    return z;
}
#else
__device__ __forceinline__ double fifth_root(double x) { return pow(x, 0.2); }
__device__ __forceinline__ double inv_tenth_root(double x) { return pow(x, -0.1); }
#endif
// 0.9 * err^(-1/5) of the step-size controller, from err2 = err^2, clamped to the range in which the
// caller's min(10, .) / max(0.2, .) can still depend on the value (outside it the clamp decides)
// A NaN error norm (non-finite state) must take the upper clamp: the factor is then below 0.2, the rejected step
// shrinks by max(0.2, .) = 0.2 as it does in SciPy (where the comparison with NaN fails), and the attempt loop
// ends at min_step with "step size too small" instead of never ending.
__device__ __forceinline__ double controller_factor(double err2) {
    const double e = (err2 <= 1e10) ? sel_max(1e-14, err2) : 1e10;
    return 0.9 * inv_tenth_root(e);
}

enum { TASK_GOAL = 0, TASK_KEPLER = 1, TASK_DNC = 2 };  // DNC: do_not_crash.py (Kepler-like scene, no orbit)
enum { STEER_ACC = 0, STEER_VEL = 1 };

// ---- Dormand-Prince tableau (scipy:rk.py:538-565), 1-based stage names ------------------
constexpr double a21 = 1.0 / 5;
constexpr double a31 = 3.0 / 40, a32 = 9.0 / 40;
constexpr double a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
constexpr double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
constexpr double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176,
                 a65 = -5103.0 / 18656;
constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
constexpr double e1 = -71.0 / 57600, e3 = 71.0 / 16695, e4 = -71.0 / 1920, e5 = 17253.0 / 339200,
                 e6 = -22.0 / 525, e7 = 1.0 / 40;
constexpr double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;

// Nystrom coefficients: position stage i = x0 + c_i h v0 + h^2 sum_k ab_ik A_k, with
// ab_ik = sum_j a_ij a_jk (the stage velocities are eliminated).
constexpr double ab31 = a32 * a21;
constexpr double ab41 = a42 * a21 + a43 * a31, ab42 = a43 * a32;
constexpr double ab51 = a52 * a21 + a53 * a31 + a54 * a41, ab52 = a53 * a32 + a54 * a42, ab53 = a54 * a43;
constexpr double ab61 = a62 * a21 + a63 * a31 + a64 * a41 + a65 * a51,
                 ab62 = a63 * a32 + a64 * a42 + a65 * a52, ab63 = a64 * a43 + a65 * a53, ab64 = a65 * a54;
constexpr double bb1 = b3 * a31 + b4 * a41 + b5 * a51 + b6 * a61, bb2 = b3 * a32 + b4 * a42 + b5 * a52 + b6 * a62,
                 bb3 = b4 * a43 + b5 * a53 + b6 * a63, bb4 = b5 * a54 + b6 * a64, bb5 = b6 * a65;
// error estimate of the position rows: h^2 sum_k eb_k A_k
constexpr double eb1 = e3 * a31 + e4 * a41 + e5 * a51 + e6 * a61 + e7 * b1,
                 eb2 = e3 * a32 + e4 * a42 + e5 * a52 + e6 * a62, eb3 = e4 * a43 + e5 * a53 + e6 * a63 + e7 * b3,
                 eb4 = e5 * a54 + e6 * a64 + e7 * b4, eb5 = e6 * a65 + e7 * b5, eb6 = e7 * b6;
// heading: theta_i = theta0 + c_i h omega0 + h^2 alpha cs_i, cs_i = sum_j a_ij c_j
constexpr double cs3 = a32 * c2, cs4 = a42 * c2 + a43 * c3, cs5 = a52 * c2 + a53 * c3 + a54 * c4,
                 cs6 = a62 * c2 + a63 * c3 + a64 * c4 + a65 * c5, cs7 = b3 * c3 + b4 * c4 + b5 * c5 + b6 * 1.0;

// dense-output matrix P (scipy:rk.py:554-565), rows = stages 1..7 (row 2 is zero)
constexpr double P11 = 1, P12 = -8048581381.0 / 2820520608.0, P13 = 8663915743.0 / 2820520608.0,
                 P14 = -12715105075.0 / 11282082432.0;
constexpr double P32 = 131558114200.0 / 32700410799.0, P33 = -68118460800.0 / 10900136933.0,
                 P34 = 87487479700.0 / 32700410799.0;
constexpr double P42 = -1754552775.0 / 470086768.0, P43 = 14199869525.0 / 1410260304.0,
                 P44 = -10690763975.0 / 1880347072.0;
constexpr double P52 = 127303824393.0 / 49829197408.0, P53 = -318862633887.0 / 49829197408.0,
                 P54 = 701980252875.0 / 199316789632.0;
constexpr double P62 = -282668133.0 / 205662961.0, P63 = 2019193451.0 / 616988883.0,
                 P64 = -1453857185.0 / 822651844.0;
constexpr double P72 = 40617522.0 / 29380423.0, P73 = -110615467.0 / 29380423.0, P74 = 69997945.0 / 29380423.0;

// ---- the hot path's constants in the constant bank ---------------------------------------------------------------
// A 64-bit literal that is not representable as a 32-bit high-word immediate costs two UMOV (or two IMAD.MOV) every
// time ptxas materialises it: the round-1 histogram of the hot kernel counted 270 UMOV + ~50 immediate moves per
// warp and step against 564 DFMA.  Read from __constant__ memory the same values arrive two per LDCU.128 in
// uniform registers, which the FP64 instructions take directly as an operand.  HT(name) selects the source;
// -DSG_NO_CONST_BANK keeps the literals (A/B builds).  Cold code (dense output, roots, resets) keeps the literals.
struct HotTab {
    double c2, c3, c4, c5, cs3, cs4, cs5, cs6, cs7;
    double c2x2, pad0;
    double ab31, ab41, ab42, ab51, ab52, ab53, ab61, ab62, ab63, ab64;
    double bb1, bb2, bb3, bb4, bb5, b1, b3, b4, b5, b6;
    double eb1, eb2, eb3, eb4, eb5, eb6, e1, e3, e4, e5, e6, e7;
    double rtol, atol, sixth, pad1;
    double ts[6];  // sin Taylor: 1/13!, -1/11!, 1/9!, -1/7!, 1/5!, -1/3!
    double tc[6];  // cos Taylor: -1/14!, 1/12!, -1/10!, 1/8!, -1/6!, 1/4!
};
static __constant__ HotTab kHot = {
    c2, c3, c4, c5, cs3, cs4, cs5, cs6, cs7,
    2 * c2, 0.0,
    ab31, ab41, ab42, ab51, ab52, ab53, ab61, ab62, ab63, ab64,
    bb1, bb2, bb3, bb4, bb5, b1, b3, b4, b5, b6,
    eb1, eb2, eb3, eb4, eb5, eb6, e1, e3, e4, e5, e6, e7,
    kRtol, kAtol, 1.0 / 6.0, 0.0,
    {1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0},
    {-1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0},
};
#ifndef SG_NO_CONST_BANK
#define HT(name) (kHot.name)
#define HTS(k) (kHot.ts[k])
#define HTC(k) (kHot.tc[k])
#else
#define HT(name) (name)
#define HTS(k) (k == 0 ? 1.0 / 6227020800.0 : k == 1 ? -1.0 / 39916800.0 : k == 2 ? 1.0 / 362880.0 : k == 3 ? -1.0 / 5040.0 : k == 4 ? 1.0 / 120.0 : -1.0 / 6.0)
#define HTC(k) (k == 0 ? -1.0 / 87178291200.0 : k == 1 ? 1.0 / 479001600.0 : k == 2 ? -1.0 / 3628800.0 : k == 3 ? 1.0 / 40320.0 : k == 4 ? -1.0 / 720.0 : 1.0 / 24.0)
#endif
constexpr double rtol = kRtol, atol = kAtol, sixth = 1.0 / 6.0;  // literal names behind HT()

// ---- engine constants, passed to kernels by value ------------------------------------------
struct DevCfg {
    // struct-of-arrays state (include/spacegym_b200.h: sg_layout)
    double* ship;
    double* planets;
    double* goal;
    double* orbit;
    double* col_shift;
    double* ep_return;
    unsigned long long* free_tiles;
    uint32_t* tile_word;
    uint32_t* elapsed;
    uint32_t* rng_ctr;
    uint32_t* status;
    double* stats;  // 4 doubles: sum return, sum length, episodes, goals
    // work queues of the split step path (library-owned scratch, empty between sg_step calls)
    uint32_t* q_counts;   // [0] continuation, [1] event, [2] post, [3] block ticket, [4..5] u64 step counter, [6] continuation 2
    uint32_t* q_cont;     // env indices, first continuation queue (filled by the hot kernel)
    uint32_t* q_cont2;    // second continuation queue (filled by the single-attempt continuation kernel)
    double* scr_cont2;
    uint32_t* q_event;
    uint32_t* q_post;     // env index | reached << 30 | done << 31
    double* scr_cont;     // [28][cap] rows, 16 in use: x y th vx vy om cth sth Ax Ay h_abs t (lx ly) rejected | fs alpha act_term | (scene)
    double* scr_event;    // [25][cap]: x y th vx vy om h t_old Ax[7] Ay[7] lx ly active
    long long q_cap;
    long long n;       // envs of this (sub-)batch
    long long stride;  // field stride of the struct-of-arrays state = envs of the whole handle
    long long first_env;
    unsigned long long seed;
    int max_episode_steps, randomize, auto_reset, rows, cols;
    int pack_flags;  // 1: done_out holds done | truncated << 1 in one byte per env (host result block), trunc_out unused
    void* term_obs;  // [N][D] rows of the step's observation dtype, or NULL: terminal observation of envs that restart in the step
#ifdef SG_CHECK_QUEUES
    // -DSG_CHECK_QUEUES (debug build, tests/test_gpu_queue_check.py): device-side invariants of the split step's work
    // queues, counted instead of trapped.  chk_err[k]: 0 queue counters not zero when a step starts, 1 queue slot >=
    // capacity, 2 an env did not leave its step exactly once through {finish, event}, 3 an env consumed twice from the
    // same queue in one step, 4 queue entry is not an env index, 5 queue counters not zero after the tail kernel.
    uint32_t* chk_err;     // [8]
    uint32_t* chk_visits;  // [N]: step exits per env since the handle was created
    uint32_t* chk_mark;    // [3][N]: epoch of the last consumption from the continuation / event / post queue
    uint32_t chk_epoch;    // steps checked so far + 1
#endif
    // physics
    double T, half_ws, two_over_ws, om_max, inv_mass, moi;
    double max_engine, max_thruster;          // float64 action path (discrete)
    float max_engine_f32, max_thruster_f32;   // float32 action path (continuous)
    double gm0;                               // G * m_ship * M_planet (all Goal planets equal)
    double gmm;                               // gm0 / m_ship: the acceleration form the force evaluations use
    double r0, r1;                            // event radii: Goal planets r0; Kepler planet r0, border r1
    // Goal reward
    double survival, goal_vel, safety, sparse, danger_zone, goal_radius, ship_radius;
    // hexagonal tiling
    double ws, hex_a, hex_height, hex_width, free_x;
    // Kepler
    double orbit_a, numC, radC, actC;
    float actC_f32;
    double kep_lo, kep_hi;  // reset radius range (kepler.py:235-237, do_not_crash.py:36-38)
    double reset_v_std, reset_om_div;  // reset velocity scale / omega divisor (kepler.py:262-265: 0.05, 5; do_not_crash.py:41-43: 0.07, 3)
    double const_reward;               // DoNotCrash: 100 / max_episode_steps (do_not_crash.py:11,47-48)
};

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) ---------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Draw session on an env's stream 0: counter = (env_lo, env_hi, block, 0); one block = 2 doubles.
struct Rng {
    uint32_t k0, k1, e0, e1, ctr;
    double ubuf, nbuf;
    bool have_u, have_n;
    __device__ __forceinline__ void open(unsigned long long seed, unsigned long long env, uint32_t ctr_) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        e0 = (uint32_t)env;
        e1 = (uint32_t)(env >> 32);
        ctr = ctr_;
        have_u = have_n = false;
    }
    __device__ __forceinline__ void fresh_block() { have_u = have_n = false; }
    // __noinline__: resets draw ~30 numbers; one copy of the Philox rounds instead of 30 keeps the cold
    // kernels' code (and their instruction-cache misses) small
    __device__ __noinline__ double uniform() {
        if (have_u) {
            have_u = false;
            return ubuf;
        }
        uint32_t o[4];
        philox4x32_10(e0, e1, ctr, 0u, k0, k1, o);
        ctr += 1;
        unsigned long long a = ((unsigned long long)o[1] << 32) | o[0];
        unsigned long long b = ((unsigned long long)o[3] << 32) | o[2];
        ubuf = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        have_u = true;
        return (double)(a >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __noinline__ double normal() {  // Box-Muller
        if (have_n) {
            have_n = false;
            return nbuf;
        }
        double u1 = 1.0 - uniform();
        double u2 = uniform();
        double r = sqrt(-2.0 * log(u1));
        double s, c;
        sincos(kTwoPi * u2, &s, &c);
        nbuf = r * s;
        have_n = true;
        return r * c;
    }
};

// synthetic actions: counter = (env_lo, env_hi, step_lo, 0x80000000 | step_hi)
__device__ __forceinline__ void random_action(unsigned long long seed, unsigned long long env,
                                              unsigned long long step, bool discrete, float& a0, float& a1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)step,
                  0x80000000u | (uint32_t)((step >> 32) & 0x7fffffffu), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    if (discrete) {
        a0 = (float)(uint32_t)(((unsigned long long)o[0] * 6u) >> 32);
        a1 = 0.0f;
    } else {
        a0 = __fsub_rn(__fmul_rn((float)(o[0] >> 8), 1.0f / 8388608.0f), 1.0f);
        a1 = __fsub_rn(__fmul_rn((float)(o[1] >> 8), 1.0f / 8388608.0f), 1.0f);
    }
}

// ---- per-thread environment registers --------------------------------------------------------
template <int TASK_, int NP_, int STEER_, int DISC_>
struct Traits {
    static constexpr int TASK = TASK_;
    static constexpr int NP = NP_;                                   // Goal: planets; Kepler: 1 gravitating body
    static constexpr int STEER = STEER_;
    static constexpr int DISC = DISC_;
    static constexpr int NEV = (TASK_ == TASK_GOAL ? NP_ : 2) + 3;   // events (dynamic_model.py:183-217)
    static constexpr int D = TASK_ == TASK_GOAL ? 9 + 2 * NP_ : (TASK_ == TASK_KEPLER ? 10 : 7);  // observation length
    static constexpr bool TILED = TASK_ == TASK_GOAL && NP_ > 1;    // hexagonal-tiling placement (one planet: goal.py:90-123)
};

template <class TR>
struct EnvRegs {
    double x, y, th, vx, vy, om;
    double px[TR::NP], py[TR::NP];  // planet centres (Kepler: the origin)
    double gx, gy;                  // goal (Goal)
    double ecc, ang;                // reference orbit (Kepler)
};

// translated action (spaceship_env.py:189-214) folded with the force scales (dynamic_model.py:140,171,175)
struct Act {
    double fs;        // engine force magnitude
    double w;         // velocity steering: heading rate
    double alpha;     // acceleration steering: angular acceleration
    double act_term;  // Kepler: act_penalty_C * ||action|| in the reference's dtype
};

template <class TR>
__device__ __forceinline__ Act translate_action(const DevCfg& c, float r0, float r1) {
    Act a;
    if (TR::DISC) {
        // DiscreteSpaceshipEnv._translate_raw_action: python floats -> float64 everywhere
        int k = (int)r0;
        double e = (k == 1 || k >= 4) ? 1.0 : 0.0;
        double t = (k == 2 || k == 4) ? -1.0 : ((k == 3 || k == 5) ? 1.0 : 0.0);
        a.fs = e * c.max_engine;
        a.w = t * 5.0;
        a.alpha = (t * c.max_thruster) / c.moi;
        a.act_term = c.actC * sqrt(e * e + t * t);
    } else {
        // float32 arithmetic, never contracted (NEP-50 promotion in the reference)
        float e32 = __fdiv_rn(__fadd_rn(r0, 1.0f), 2.0f);
        float t32 = r1;
        a.fs = (double)__fmul_rn(e32, c.max_engine_f32);
        a.w = (double)__fmul_rn(t32, 5.0f);
        a.alpha = (double)__fmul_rn(t32, c.max_thruster_f32) / c.moi;
        float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(e32, e32), __fmul_rn(t32, t32)));
        a.act_term = (double)__fmul_rn(c.actC_f32, nrm);
    }
    return a;
}

// ---- forces (dynamic_model.py:145-176, helpers.py:22-35) -----------------------------------------
// a = (-(cos th, sin th) fs + sum_p GM d_p / |d_p|^3) / m.  The Kepler border has mass 0 and
// contributes exactly 0, so it is not a gravitating body here.
// |d|^-3 from d2 = |d|^2: hardware rsqrt seed y (~2^-22) and ONE correction of the -3/2 power,
// d2^(-3/2) = y^3 (1 - e)^(-3/2) = y^3 (1 + 3/2 e + 15/8 e^2 + O(e^3)), e = 1 - d2 y^2 (|e| < 5e-7: the dropped term is
// < 3e-19).  The force only needs the cube, so this replaces rsqrt-to-full-precision followed by two multiplies: the
// dependent chain behind the MUFU is 4 FP64 operations instead of 7.  Built with -DSG_SHORT_CHAIN only: measured
// neutral (5.735e9 vs 5.726e9 env-steps/s, profiles/r2_notes.md) — four warps per scheduler already saturate the
// FP64 pipe on dependent DFMA chains (tools/micro/fp64_latency.cu: 8-cycle latency, one DFMA per 2 cycles), so the
// chain length is not what the hot kernel waits for.  The round-1 form stays the default.
__device__ __forceinline__ double inv_cube(double d2) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
    const double y2 = y * y;
    const double e = fma(-d2, y2, 1.0);
    const double y3 = y2 * y;
    const double p = fma(e, 1.875, 1.5);
    return fma(y3 * e, p, y3);
}

template <class TR>
__device__ __forceinline__ void accel(const DevCfg& c, const EnvRegs<TR>& e, double x, double y, double cth,
                                      double sth, double fs, double& ax, double& ay) {
#ifndef SG_SHORT_CHAIN
    double fx = -cth * fs, fy = -sth * fs;
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        double dx = e.px[p] - x, dy = e.py[p] - y;
        double inv = fast_rsqrt(dx * dx + dy * dy);
        double s = c.gm0 * (inv * inv * inv);
        fx = fma(dx, s, fx);
        fy = fma(dy, s, fy);
    }
    ax = fx * c.inv_mass;
    ay = fy * c.inv_mass;
#else
    // thrust and GM already divided by the ship's mass (both products are common to every force evaluation of the
    // step: computed once); the planets' terms accumulate onto the thrust term
    const double fsm = fs * c.inv_mass;
    double fx = -cth * fsm, fy = -sth * fsm;
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        const double dx = e.px[p] - x, dy = e.py[p] - y;
        const double s = c.gmm * inv_cube(fma(dx, dx, dy * dy));
        fx = fma(dx, s, fx);
        fy = fma(dy, s, fy);
    }
    ax = fx;
    ay = fy;
#endif
}

template <class TR>
__device__ __forceinline__ void event_values_from_pd(const DevCfg& c, double x, double y, double om,
                                                     const double (&pd)[TR::NP], double (&g)[TR::NEV]);

// same force, also returning |planet - ship| and its inverse: the event check at the same point reuses them
template <class TR>
__device__ __forceinline__ void accel_pd(const DevCfg& c, const EnvRegs<TR>& e, double x, double y, double cth,
                                         double sth, double fs, double& ax, double& ay, double (&pd)[TR::NP],
                                         double (&ipd)[TR::NP]) {
#ifndef SG_SHORT_CHAIN
    double fx = -cth * fs, fy = -sth * fs;
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        double dx = e.px[p] - x, dy = e.py[p] - y;
        const double d2 = dx * dx + dy * dy;
        const double inv = fast_rsqrt(d2);
        const double sq = d2 * inv;
        pd[p] = fma(fma(-sq, sq, d2), 0.5 * inv, sq);  // fast_sqrt(d2) from the same rsqrt
        ipd[p] = inv;
        double s = c.gm0 * (inv * inv * inv);
        fx = fma(dx, s, fx);
        fy = fma(dy, s, fy);
    }
    ax = fx * c.inv_mass;
    ay = fy * c.inv_mass;
#else
    const double fsm = fs * c.inv_mass;
    double fx = -cth * fsm, fy = -sth * fsm;
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        const double dx = e.px[p] - x, dy = e.py[p] - y;
        const double d2 = fma(dx, dx, dy * dy);
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));
        const double y2 = y0 * y0;
        const double er = fma(-d2, y2, 1.0);
        const double y3 = y2 * y0;
        // force: the cube directly (as inv_cube); distance and its inverse from the same seed, off the force's chain
        const double s = c.gmm * fma(y3 * er, fma(er, 1.875, 1.5), y3);
        fx = fma(dx, s, fx);
        fy = fma(dy, s, fy);
        const double inv = fma(y0 * er, fma(er, 0.375, 0.5), y0);
        const double sq = d2 * inv;
        pd[p] = fma(fma(-sq, sq, d2), 0.5 * inv, sq);  // fast_sqrt(d2)
        ipd[p] = inv;
    }
    ax = fx;
    ay = fy;
#endif
}

// ---- termination events (dynamic_model.py:183-217) -------------------------------------------------
// order: planets..., world_max, world_min, angular velocity.  pd = planet-centre distances.
template <class TR>
__device__ __forceinline__ void event_values(const DevCfg& c, const EnvRegs<TR>& e, double x, double y, double om,
                                             double (&g)[TR::NEV], double (&pd)[TR::NP]) {
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        double dx = e.px[p] - x, dy = e.py[p] - y;
        pd[p] = fast_sqrt(dx * dx + dy * dy);
    }
    event_values_from_pd<TR>(c, x, y, om, pd, g);
}

// same, also returning 1 / distance (the lidar rows of the observation reuse it)
template <class TR>
__device__ __forceinline__ void event_values_inv(const DevCfg& c, const EnvRegs<TR>& e, double x, double y, double om,
                                                 double (&g)[TR::NEV], double (&pd)[TR::NP], double (&ipd)[TR::NP]) {
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        double dx = e.px[p] - x, dy = e.py[p] - y;
        const double d2 = dx * dx + dy * dy;
        const double inv = fast_rsqrt(d2);
        const double s = d2 * inv;
        pd[p] = fma(fma(-s, s, d2), 0.5 * inv, s);
        ipd[p] = inv;
    }
    event_values_from_pd<TR>(c, x, y, om, pd, g);
}

template <class TR>
__device__ __forceinline__ void event_values_from_pd(const DevCfg& c, double x, double y, double om,
                                                     const double (&pd)[TR::NP], double (&g)[TR::NEV]) {
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) g[p] = pd[p] - c.r0;
    } else {
        g[0] = pd[0] - c.r0;
        g[1] = pd[0] - c.r1;
    }
    g[TR::NEV - 3] = sel_min(c.half_ws - x, c.half_ws - y);
    g[TR::NEV - 2] = sel_min(c.half_ws + x, c.half_ws + y);
    g[TR::NEV - 1] = c.om_max - fabs(om);
}

// ---- cold path: dense output + Brent roots of the active events -------------------------------------
template <class TR>
struct EventIO {
    // in: state at the start of the accepted step, the step, its seven force evaluations
    double x, y, th, vx, vy, om, alpha;
    double h, t_old;
    double Ax[7], Ay[7];
    double px[TR::NP], py[TR::NP];
    uint32_t active;
    // out
    double t_root;
    uint32_t status;
};

template <class TR>
struct Dense {  // scipy:rk.py:715-737, specialised: heading and its rate are exact polynomials
    double x0, y0, th0, vx0, vy0, om0, alpha, h, t_old;
    double Qx[4], Qy[4], Qvx[4], Qvy[4];
    __device__ void eval(double t, double& x, double& y, double& th, double& vx, double& vy, double& om) const {
        double s = (t - t_old) / h;
        double s2 = s * s, s3 = s2 * s, s4 = s3 * s;
        x = x0 + h * (Qx[0] * s + Qx[1] * s2 + Qx[2] * s3 + Qx[3] * s4);
        y = y0 + h * (Qy[0] * s + Qy[1] * s2 + Qy[2] * s3 + Qy[3] * s4);
        vx = vx0 + h * (Qvx[0] * s + Qvx[1] * s2 + Qvx[2] * s3 + Qvx[3] * s4);
        vy = vy0 + h * (Qvy[0] * s + Qvy[1] * s2 + Qvy[2] * s3 + Qvy[3] * s4);
        double dt = t - t_old;
        om = om0 + alpha * dt;
        th = th0 + om0 * dt + 0.5 * alpha * dt * dt;
    }
    // what the event functions look at: position and heading rate
    __device__ void eval_xyo(double t, double& x, double& y, double& om) const {
        double s = (t - t_old) / h;
        double s2 = s * s, s3 = s2 * s, s4 = s3 * s;
        x = x0 + h * (Qx[0] * s + Qx[1] * s2 + Qx[2] * s3 + Qx[3] * s4);
        y = y0 + h * (Qy[0] * s + Qy[1] * s2 + Qy[2] * s3 + Qy[3] * s4);
        om = om0 + alpha * (t - t_old);
    }
};

// value of event `ev` at time t of the accepted step; `ev` may differ from lane to lane, so the selection of
// the planet and of the event kind is done with selects instead of indexed (local-memory) accesses
template <class TR>
__device__ double event_at(const DevCfg& c, const EventIO<TR>& io, const Dense<TR>& d, int ev, double t) {
    double x, y, om;
    d.eval_xyo(t, x, y, om);
    constexpr int NPL = TR::NEV - 3;  // planet-type events
    double pxs = io.px[0], pys = io.py[0];
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 1; p < TR::NP; ++p) {
            pxs = (ev == p) ? io.px[p] : pxs;
            pys = (ev == p) ? io.py[p] : pys;
        }
    }
    const double r = (TR::TASK == TASK_GOAL) ? c.r0 : (ev == 0 ? c.r0 : c.r1);
    const double dx = pxs - x, dy = pys - y;
    const double g_planet = sqrt(dx * dx + dy * dy) - r;
    const double g_max = fmin(c.half_ws - x, c.half_ws - y);
    const double g_min = fmin(c.half_ws + x, c.half_ws + y);
    const double g_om = c.om_max - fabs(om);
    return ev < NPL ? g_planet : (ev == NPL ? g_max : (ev == NPL + 1 ? g_min : g_om));
}

// Brent's method as scipy.optimize.brentq runs it from scipy:ivp.py:52-77 (xtol = rtol = 4 eps,
// 100 iterations): bisection guarded secant / inverse quadratic interpolation.
template <class TR>
__device__ double brent_root(const DevCfg& c, const EventIO<TR>& io, const Dense<TR>& d, int ev, double xa,
                             double xb, uint32_t& status) {
    const double tol = 4 * kEps;
    double xpre = xa, xcur = xb, xblk = 0, fblk = 0, spre = 0, scur = 0;
    double fpre = event_at(c, io, d, ev, xpre), fcur = event_at(c, io, d, ev, xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) {
        status |= 2u;  // SG_ENV_BRENT_SIGN: the reference would raise here
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
        double delta = (tol + tol * fabs(xcur)) / 2;
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
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
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
        fcur = event_at(c, io, d, ev, xcur);
    }
    return xcur;
}

// scipy:ivp.py:683-697: roots of all active events, earliest one terminates; state := sol(t*)
template <class TR>
static __device__ __noinline__ void handle_events(const DevCfg& c, EventIO<TR>* io) {
    Dense<TR> d;
    d.x0 = io->x; d.y0 = io->y; d.th0 = io->th; d.vx0 = io->vx; d.vy0 = io->vy; d.om0 = io->om;
    d.alpha = io->alpha; d.h = io->h; d.t_old = io->t_old;
    const double h = io->h;
    const double* Ax = io->Ax;
    const double* Ay = io->Ay;
    // stage velocities (rows 0-1 of K): V_i = v + h sum_j a_ij A_j
    double Vx[7], Vy[7];
    Vx[0] = io->vx; Vy[0] = io->vy;
    Vx[1] = io->vx + h * (a21 * Ax[0]);
    Vy[1] = io->vy + h * (a21 * Ay[0]);
    Vx[2] = io->vx + h * (a31 * Ax[0] + a32 * Ax[1]);
    Vy[2] = io->vy + h * (a31 * Ay[0] + a32 * Ay[1]);
    Vx[3] = io->vx + h * (a41 * Ax[0] + a42 * Ax[1] + a43 * Ax[2]);
    Vy[3] = io->vy + h * (a41 * Ay[0] + a42 * Ay[1] + a43 * Ay[2]);
    Vx[4] = io->vx + h * (a51 * Ax[0] + a52 * Ax[1] + a53 * Ax[2] + a54 * Ax[3]);
    Vy[4] = io->vy + h * (a51 * Ay[0] + a52 * Ay[1] + a53 * Ay[2] + a54 * Ay[3]);
    Vx[5] = io->vx + h * (a61 * Ax[0] + a62 * Ax[1] + a63 * Ax[2] + a64 * Ax[3] + a65 * Ax[4]);
    Vy[5] = io->vy + h * (a61 * Ay[0] + a62 * Ay[1] + a63 * Ay[2] + a64 * Ay[3] + a65 * Ay[4]);
    Vx[6] = io->vx + h * (b1 * Ax[0] + b3 * Ax[2] + b4 * Ax[3] + b5 * Ax[4] + b6 * Ax[5]);
    Vy[6] = io->vy + h * (b1 * Ay[0] + b3 * Ay[2] + b4 * Ay[3] + b5 * Ay[4] + b6 * Ay[5]);
    const double Pm[7][4] = {{P11, P12, P13, P14}, {0, 0, 0, 0},       {0, P32, P33, P34}, {0, P42, P43, P44},
                             {0, P52, P53, P54},   {0, P62, P63, P64}, {0, P72, P73, P74}};
    for (int j = 0; j < 4; ++j) {
        double qx = 0, qy = 0, qvx = 0, qvy = 0;
        for (int k = 0; k < 7; ++k) {
            qx += Vx[k] * Pm[k][j];
            qy += Vy[k] * Pm[k][j];
            qvx += Ax[k] * Pm[k][j];
            qvy += Ay[k] * Pm[k][j];
        }
        d.Qx[j] = qx; d.Qy[j] = qy; d.Qvx[j] = qvx; d.Qvy[j] = qvy;
    }
    double best = INFINITY;
    uint32_t status = 0;
    // every lane walks its OWN active events (usually one): the warp makes max-popcount passes through Brent's
    // method instead of one pass per event kind that any lane has active
    for (uint32_t mask = io->active; mask; mask &= mask - 1u) {
        const int ev = __ffs((int)mask) - 1;
        double root = brent_root(c, *io, d, ev, io->t_old, io->t_old + io->h, status);
        if (root < best) best = root;
    }
    io->t_root = best;
    io->status = status;
    d.eval(best, io->x, io->y, io->th, io->vx, io->vy, io->om);
}

// ---- integrator state carried between accepted steps ------------------------------------------------
template <class TR>
struct Integ {
    double x, y, th, vx, vy, om;  // current state (omega already overwritten in velocity mode)
    double cth, sth;              // cos / sin of th
    double Ax, Ay;                // force evaluation at the current state (FSAL slot, K[0] of the next step)
    double h_abs, t;              // solver.h_abs, solver.t
    double alpha, fs;             // constant over the env step
};

// cos / sin of a small heading increment by Taylor polynomials (|d| < 0.5: error < 3e-17), so that the
// stage headings cost a rotation of (cos th, sin th) instead of a full-range sincos each.
__device__ __forceinline__ void small_sincos(double d, bool small, double& sn, double& cs) {
    if (small) {
        const double d2 = d * d;
        double ps = fma(d2, HTS(0), HTS(1));
        double pc = fma(d2, HTC(0), HTC(1));
#pragma unroll
        for (int k = 2; k < 6; ++k) {
            ps = fma(ps, d2, HTS(k));
            pc = fma(pc, d2, HTC(k));
        }
        sn = fma(ps * d2, d, d);
        pc = fma(pc, d2, -0.5);
        cs = fma(pc, d2, 1.0);
    } else {
        sincos(d, &sn, &cs);
    }
}

// force at (x, y) with the heading rotated by d from the step's base heading
template <class TR>
__device__ __forceinline__ void accel_rot(const DevCfg& c, const EnvRegs<TR>& e, const Integ<TR>& S, double x,
                                          double y, double d, bool small, double& ax, double& ay, double& cn,
                                          double& sn) {
    double sd, cd;
    small_sincos(d, small, sd, cd);
    cn = S.cth * cd - S.sth * sd;
    sn = S.sth * cd + S.cth * sd;
    accel(c, e, x, y, cn, sn, S.fs, ax, ay);
}

// RK45.__init__ (scipy:rk.py:85-103): f0, select_initial_step (scipy:common.py:109-134, error estimator
// order 4), and the event values at t0 (scipy:ivp.py:651, after the omega overwrite of quirk 1).
template <class TR>
__device__ __forceinline__ void init_step(const DevCfg& c, EnvRegs<TR>& e, const Act& act, Integ<TR>& S,
                                          double (&g)[TR::NEV], double (&pd)[TR::NP]) {
    const double T = c.T;
    S.alpha = (TR::STEER == STEER_ACC) ? act.alpha : 0.0;
    S.fs = act.fs;
    if (TR::STEER == STEER_VEL) e.om = act.w;  // the RHS overwrites omega in place
    S.x = e.x; S.y = e.y; S.th = e.th; S.vx = e.vx; S.vy = e.vy; S.om = e.om;
    S.t = 0.0;
#if SG_FAST_SINCOS0
    // the heading of a stepped env is wrapped to [0, 2 pi) (wrap_ship_angle): quadrant reduction + Taylor, no slow
    // path; states injected from outside the range take the library routine
    if (S.th >= 0.0 && S.th <= kTwoPi) sincos_2pi(S.th, S.sth, S.cth);
    else sincos_cold(S.th, &S.sth, &S.cth);
#else
    sincos(S.th, &S.sth, &S.cth);
#endif
    double ipd0[TR::NP];
    accel_pd<TR>(c, e, S.x, S.y, S.cth, S.sth, S.fs, S.Ax, S.Ay, pd, ipd0);  // f0 (+ planet distances for g(t0))
    const double x = S.x, y = S.y, th = S.th, vx = S.vx, vy = S.vy, om = S.om, alpha = S.alpha;
    double rx = fast_rcp(HT(atol) + fabs(x) * HT(rtol)), ry = fast_rcp(HT(atol) + fabs(y) * HT(rtol));
    double rth = fast_rcp(HT(atol) + fabs(th) * HT(rtol)), rvx = fast_rcp(HT(atol) + fabs(vx) * HT(rtol));
    double rvy = fast_rcp(HT(atol) + fabs(vy) * HT(rtol)), rom = fast_rcp(HT(atol) + fabs(om) * HT(rtol));
    double q, S0 = 0, S1 = 0, S2 = 0;
    q = x * rx; S0 = fma(q, q, S0);
    q = y * ry; S0 = fma(q, q, S0);
    q = th * rth; S0 = fma(q, q, S0);
    q = vx * rvx; S0 = fma(q, q, S0);
    q = vy * rvy; S0 = fma(q, q, S0);
    q = om * rom; S0 = fma(q, q, S0);
    q = vx * rx; S1 = fma(q, q, S1);
    q = vy * ry; S1 = fma(q, q, S1);
    q = om * rth; S1 = fma(q, q, S1);
    q = S.Ax * rvx; S1 = fma(q, q, S1);
    q = S.Ay * rvy; S1 = fma(q, q, S1);
    q = alpha * rom; S1 = fma(q, q, S1);
    // select_initial_step on the SQUARES of its norms (d0^2 = S0 / 6, d1^2 = S1 / 6, d2^2 = S2 / (6 h0^2)): every
    // comparison of scipy:common.py:118-132 is monotone in the norm, h0 = 0.01 d0 / d1 = 0.01 sqrt(S0 / S1) needs
    // one reciprocal and one square root instead of two square roots and a division, and d2 is never divided out.
    double h0 = (S0 < 6e-10 || S1 < 6e-10) ? 1e-6 : 0.01 * fast_sqrt(S0 * fast_rcp(S1));
    h0 = sel_min(T, h0);
    double Bx, By, cn, sn;
    const double dprobe = h0 * om;
    accel_rot(c, e, S, x + h0 * vx, y + h0 * vy, dprobe, fabs(dprobe) < 0.5, Bx, By, cn, sn);  // f(y0 + h0 f0)
    q = h0 * S.Ax * rx; S2 = fma(q, q, S2);
    q = h0 * S.Ay * ry; S2 = fma(q, q, S2);
    q = h0 * alpha * rth; S2 = fma(q, q, S2);
    q = (Bx - S.Ax) * rvx; S2 = fma(q, q, S2);
    q = (By - S.Ay) * rvy; S2 = fma(q, q, S2);
    const double hq = 6.0 * (h0 * h0);
    const double M = sel_max(S1 * (h0 * h0), S2);  // = 6 h0^2 max(d1, d2)^2
    const double lim = sel_min(T, 100 * h0);
    double h_abs;
    if (S1 <= 6e-30 && S2 <= 1e-30 * hq) {  // d1 <= 1e-15 and d2 <= 1e-15
        h_abs = sel_min(lim, sel_max(1e-6, h0 * 1e-3));
    } else {
        // h1 = (0.01 / dm) ** 0.2 only matters when it is below lim: decide that without the root
        // (0.01 >= dm lim^5  <=>  1e-4 * 6 h0^2 >= M lim^10), and pay for it only on the lanes that need its value:
        // h1 = (dm^2 / 1e-4) ** (-1/10).
        const double l2 = lim * lim, l5 = l2 * l2 * lim;
        if (1e-4 * hq >= M * (l5 * l5) * 1.0000000002) h_abs = lim;
        else h_abs = sel_min(lim, inv_tenth_root(sel_min(1e30, sel_max(1e-30, fast_div(1e4 * M, hq)))));
    }
    S.h_abs = h_abs;
    event_values_from_pd<TR>(c, x, y, om, pd, g);
}


struct SinCos6 {
    double s[6], c[6];
};
static __device__ __noinline__ SinCos6 slow_sincos6(double d0, double d1, double d2, double d3, double d4, double d5) {
    SinCos6 r;
    const double d[6] = {d0, d1, d2, d3, d4, d5};
#pragma unroll 1
    for (int j = 0; j < 6; ++j) sincos(d[j], &r.s[j], &r.c[j]);
    return r;
}

// cos / sin of N small heading increments at once: the Horner chains are interleaved (N-way ILP) and every
// Taylor coefficient is materialised once for all N.
template <int N>
__device__ __forceinline__ void small_sincos_n(const double (&d)[N], bool small, double (&sn)[N], double (&cs)[N]) {
    if (small) {
        double d2[N], ps[N], pc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            d2[j] = d[j] * d[j];
            ps[j] = fma(d2[j], HTS(0), HTS(1));
            pc[j] = fma(d2[j], HTC(0), HTC(1));
        }
#pragma unroll
        for (int k = 2; k < 6; ++k) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                ps[j] = fma(ps[j], d2[j], HTS(k));
                pc[j] = fma(pc[j], d2[j], HTC(k));
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            sn[j] = fma(ps[j] * d2[j], d[j], d[j]);
            pc[j] = fma(pc[j], d2[j], -0.5);
            cs[j] = fma(pc[j], d2[j], 1.0);
        }
    } else if (N == 6) {  // |increment| >= 0.5 rad: only states injected from outside (|omega| > omega_max)
        const SinCos6 r = slow_sincos6(d[0], d[1 % N], d[2 % N], d[3 % N], d[4 % N], d[5 % N]);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            sn[j] = r.s[j];
            cs[j] = r.c[j];
        }
    } else {  // shorter batches are only ever called with small == true
#pragma unroll 1
        for (int j = 0; j < N; ++j) sincos_cold(d[j], &sn[j], &cs[j]);
    }
}

// One Runge-Kutta attempt of size h from S (rk_step, scipy:rk.py:14-71) in Nystrom form, plus the error
// norm (scipy:rk.py:105-109,141-143; the heading rows integrate exactly and contribute 0).  The six stage
// headings depend only on (h, omega, alpha), so their cos / sin are computed up front, together.
template <class TR>
__device__ __forceinline__ void rk_attempt(const DevCfg& c, const EnvRegs<TR>& e, const Integ<TR>& S, double h,
                                           double (&Ax)[7], double (&Ay)[7], double& xn, double& yn, double& thn,
                                           double& vxn, double& vyn, double& omn, double& cn, double& sn,
                                           double& err2, double (&pdn)[TR::NP], double (&ipdn)[TR::NP]) {
    const double x = S.x, y = S.y, vx = S.vx, vy = S.vy;
    const double hh = h * h, ho = h * S.om, ha = hh * S.alpha;
    const bool small = fabs(ho) + fabs(ha) < 0.5;
    double d[6], sd[6], cd[6];
    d[0] = HT(c2) * ho;
    d[1] = HT(c3) * ho + HT(cs3) * ha;
    d[2] = HT(c4) * ho + HT(cs4) * ha;
    d[3] = HT(c5) * ho + HT(cs5) * ha;
    d[4] = ho + HT(cs6) * ha;
    d[5] = ho + HT(cs7) * ha;
    if (TR::STEER == STEER_VEL && small) {
        // velocity steering: alpha = 0, so the increments are c_i h omega with c = 1/5, 3/10, 4/5, 8/9, 1, 1.
        // Three polynomials (a = d[0], d[1], d[3]) instead of six: 4a by two angle doublings, 5a = 4a + a by one
        // rotation, and the last two headings coincide (d[4] == d[5] bit for bit: c4 == 4 c2 exactly and ha == 0).
        const double dd[3] = {d[0], d[1], d[3]};
        double s3[3], c3[3];
        small_sincos_n<3>(dd, true, s3, c3);
        sd[0] = s3[0]; cd[0] = c3[0];
        sd[1] = s3[1]; cd[1] = c3[1];
        sd[3] = s3[2]; cd[3] = c3[2];
        const double s2 = 2.0 * (s3[0] * c3[0]), c2a = fma(-2.0 * s3[0], s3[0], 1.0);  // sin 2a, cos 2a
        sd[2] = 2.0 * (s2 * c2a);
        cd[2] = fma(-2.0 * s2, s2, 1.0);
        sd[4] = fma(sd[2], c3[0], cd[2] * s3[0]);
        cd[4] = fma(cd[2], c3[0], -(sd[2] * s3[0]));
        sd[5] = sd[4];
        cd[5] = cd[4];
    } else {
        small_sincos_n<6>(d, small, sd, cd);
    }
    double cr[6], sr[6];  // stage headings: the base heading rotated by d[j]
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        cr[j] = S.cth * cd[j] - S.sth * sd[j];
        sr[j] = S.sth * cd[j] + S.cth * sd[j];
    }
    Ax[0] = S.Ax;
    Ay[0] = S.Ay;
#ifndef SG_SHORT_CHAIN
    accel(c, e, x + (HT(c2) * h) * vx, y + (HT(c2) * h) * vy, cr[0], sr[0], S.fs, Ax[1], Ay[1]);
    accel(c, e, x + (HT(c3) * h) * vx + hh * (HT(ab31) * Ax[0]), y + (HT(c3) * h) * vy + hh * (HT(ab31) * Ay[0]), cr[1], sr[1], S.fs,
          Ax[2], Ay[2]);
    accel(c, e, x + (HT(c4) * h) * vx + hh * (HT(ab41) * Ax[0] + HT(ab42) * Ax[1]),
          y + (HT(c4) * h) * vy + hh * (HT(ab41) * Ay[0] + HT(ab42) * Ay[1]), cr[2], sr[2], S.fs, Ax[3], Ay[3]);
    accel(c, e, x + (HT(c5) * h) * vx + hh * (HT(ab51) * Ax[0] + HT(ab52) * Ax[1] + HT(ab53) * Ax[2]),
          y + (HT(c5) * h) * vy + hh * (HT(ab51) * Ay[0] + HT(ab52) * Ay[1] + HT(ab53) * Ay[2]), cr[3], sr[3], S.fs, Ax[4], Ay[4]);
    accel(c, e, x + h * vx + hh * (HT(ab61) * Ax[0] + HT(ab62) * Ax[1] + HT(ab63) * Ax[2] + HT(ab64) * Ax[3]),
          y + h * vy + hh * (HT(ab61) * Ay[0] + HT(ab62) * Ay[1] + HT(ab63) * Ay[2] + HT(ab64) * Ay[3]), cr[4], sr[4], S.fs, Ax[5],
          Ay[5]);
    xn = x + h * vx + hh * (HT(bb1) * Ax[0] + HT(bb2) * Ax[1] + HT(bb3) * Ax[2] + HT(bb4) * Ax[3] + HT(bb5) * Ax[4]);
    yn = y + h * vy + hh * (HT(bb1) * Ay[0] + HT(bb2) * Ay[1] + HT(bb3) * Ay[2] + HT(bb4) * Ay[3] + HT(bb5) * Ay[4]);
    vxn = vx + h * (HT(b1) * Ax[0] + HT(b3) * Ax[2] + HT(b4) * Ax[3] + HT(b5) * Ax[4] + HT(b6) * Ax[5]);
    vyn = vy + h * (HT(b1) * Ay[0] + HT(b3) * Ay[2] + HT(b4) * Ay[3] + HT(b5) * Ay[4] + HT(b6) * Ay[5]);
#else
    // Stage positions x_i = x + c_i h vx + h^2 sum_k ab_ik A_k, accumulated term by term with the NEWEST force last:
    // everything but the final fma of a stage depends only on older forces and is off the critical path
    // (h^2 ab_ik are formed once per attempt).  One fma then separates A_{i-1} from the position it is needed at.
    {
        const double bx1 = fma(HT(c2) * h, vx, x), by1 = fma(HT(c2) * h, vy, y);
        accel(c, e, bx1, by1, cr[0], sr[0], S.fs, Ax[1], Ay[1]);
        const double h31 = hh * HT(ab31);
        const double h41 = hh * HT(ab41), h42 = hh * HT(ab42);
        const double h51 = hh * HT(ab51), h52 = hh * HT(ab52), h53 = hh * HT(ab53);
        const double h61 = hh * HT(ab61), h62 = hh * HT(ab62), h63 = hh * HT(ab63), h64 = hh * HT(ab64);
        const double g1 = hh * HT(bb1), g2 = hh * HT(bb2), g3 = hh * HT(bb3), g4 = hh * HT(bb4), g5 = hh * HT(bb5);
        const double x3 = fma(h31, Ax[0], fma(HT(c3) * h, vx, x)), y3 = fma(h31, Ay[0], fma(HT(c3) * h, vy, y));
        // partial sums over forces that already exist
        double x4 = fma(h41, Ax[0], fma(HT(c4) * h, vx, x)), y4 = fma(h41, Ay[0], fma(HT(c4) * h, vy, y));
        double x5 = fma(h51, Ax[0], fma(HT(c5) * h, vx, x)), y5 = fma(h51, Ay[0], fma(HT(c5) * h, vy, y));
        const double xe = fma(h, vx, x), ye = fma(h, vy, y);
        double x6 = fma(h61, Ax[0], xe), y6 = fma(h61, Ay[0], ye);
        double x7 = fma(g1, Ax[0], xe), y7 = fma(g1, Ay[0], ye);
        x4 = fma(h42, Ax[1], x4); y4 = fma(h42, Ay[1], y4);
        x5 = fma(h52, Ax[1], x5); y5 = fma(h52, Ay[1], y5);
        x6 = fma(h62, Ax[1], x6); y6 = fma(h62, Ay[1], y6);
        x7 = fma(g2, Ax[1], x7); y7 = fma(g2, Ay[1], y7);
        accel(c, e, x3, y3, cr[1], sr[1], S.fs, Ax[2], Ay[2]);
        accel(c, e, x4, y4, cr[2], sr[2], S.fs, Ax[3], Ay[3]);
        x5 = fma(h53, Ax[2], x5); y5 = fma(h53, Ay[2], y5);
        x6 = fma(h63, Ax[2], x6); y6 = fma(h63, Ay[2], y6);
        x7 = fma(g3, Ax[2], x7); y7 = fma(g3, Ay[2], y7);
        accel(c, e, x5, y5, cr[3], sr[3], S.fs, Ax[4], Ay[4]);
        x6 = fma(h64, Ax[3], x6); y6 = fma(h64, Ay[3], y6);
        x7 = fma(g4, Ax[3], x7); y7 = fma(g4, Ay[3], y7);
        accel(c, e, x6, y6, cr[4], sr[4], S.fs, Ax[5], Ay[5]);
        xn = fma(g5, Ax[4], x7);
        yn = fma(g5, Ay[4], y7);
        const double k1 = h * HT(b1), k3 = h * HT(b3), k4 = h * HT(b4), k5 = h * HT(b5), k6 = h * HT(b6);
        vxn = fma(k6, Ax[5], fma(k5, Ax[4], fma(k4, Ax[3], fma(k3, Ax[2], fma(k1, Ax[0], vx)))));
        vyn = fma(k6, Ay[5], fma(k5, Ay[4], fma(k4, Ay[3], fma(k3, Ay[2], fma(k1, Ay[0], vy)))));
    }
#endif
    thn = S.th + d[5];
    omn = S.om + h * S.alpha;
    cn = cr[5];
    sn = sr[5];
    accel_pd<TR>(c, e, xn, yn, cn, sn, S.fs, Ax[6], Ay[6], pdn, ipdn);  // f_new (FSAL) + distances at the end point
    double ex = hh * (HT(eb1) * Ax[0] + HT(eb2) * Ax[1] + HT(eb3) * Ax[2] + HT(eb4) * Ax[3] + HT(eb5) * Ax[4] + HT(eb6) * Ax[5]);
    double ey = hh * (HT(eb1) * Ay[0] + HT(eb2) * Ay[1] + HT(eb3) * Ay[2] + HT(eb4) * Ay[3] + HT(eb5) * Ay[4] + HT(eb6) * Ay[5]);
    double evx = h * (HT(e1) * Ax[0] + HT(e3) * Ax[2] + HT(e4) * Ax[3] + HT(e5) * Ax[4] + HT(e6) * Ax[5] + HT(e7) * Ax[6]);
    double evy = h * (HT(e1) * Ay[0] + HT(e3) * Ay[2] + HT(e4) * Ay[3] + HT(e5) * Ay[4] + HT(e6) * Ay[5] + HT(e7) * Ay[6]);
    ex *= fast_rcp(HT(atol) + sel_max(fabs(x), fabs(xn)) * HT(rtol));
    ey *= fast_rcp(HT(atol) + sel_max(fabs(y), fabs(yn)) * HT(rtol));
    evx *= fast_rcp(HT(atol) + sel_max(fabs(vx), fabs(vxn)) * HT(rtol));
    evy *= fast_rcp(HT(atol) + sel_max(fabs(vy), fabs(vyn)) * HT(rtol));
    err2 = (ex * ex + ey * ey + evx * evx + evy * evy) * HT(sixth);  // error_norm ** 2
}

__device__ __forceinline__ double min_step_at(double t) {  // 10 * |nextafter(t, inf) - t|, scipy:rk.py:119
    return 10 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
}

// step-size controller after an attempt (scipy:rk.py:145-165); returns true when the step is accepted
// err2 = error_norm ** 2: `error_norm < 1` and `error_norm ** -0.2` are evaluated on the square
// (x < 1 <=> x*x < 1 for x >= 0; err ** -0.2 = err2 ** -0.1), which saves the square root.
__device__ __forceinline__ bool controller(double err2, bool& rejected, bool more_steps, double& h_abs) {
    if (err2 < 1) {
        if (more_steps) {  // the next step size only matters if another step follows
            double factor = (err2 == 0) ? 10.0 : sel_min(10.0, controller_factor(err2));
            if (rejected) factor = sel_min(1.0, factor);
            h_abs *= factor;
        }
        return true;
    }
    h_abs *= sel_max(0.2, controller_factor(err2));
    rejected = true;
    return false;
}

template <class TR>
__device__ __forceinline__ uint32_t active_events(const double (&g)[TR::NEV], const double (&gn)[TR::NEV]) {
    uint32_t active = 0;  // find_active_events, scipy:ivp.py:134-157 (direction 0)
#pragma unroll
    for (int i = 0; i < TR::NEV; ++i) {
        bool up = (g[i] <= 0) && (gn[i] >= 0);
        bool down = (g[i] >= 0) && (gn[i] <= 0);
        if (up || down) active |= (1u << i);
    }
    return active;
}

template <class TR>
__device__ __forceinline__ void fill_event_io(EventIO<TR>& io, const EnvRegs<TR>& e, const Integ<TR>& S, double h,
                                              const double (&Ax)[7], const double (&Ay)[7], uint32_t active) {
    io.x = S.x; io.y = S.y; io.th = S.th; io.vx = S.vx; io.vy = S.vy; io.om = S.om; io.alpha = S.alpha;
    io.h = h; io.t_old = S.t;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        io.Ax[k] = Ax[k];
        io.Ay[k] = Ay[k];
    }
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        io.px[p] = e.px[p];
        io.py[p] = e.py[p];
    }
    io.active = active;
}

enum { RK_FINISHED = 0, RK_EVENT = 1, RK_FAILED = 2 };

// solve_ivp's loop (scipy:ivp.py:659-699 around RungeKutta._step_impl, scipy:rk.py:111-176) from the
// integrator state S at time S.t.  `rejected` is the step_rejected flag of the attempt loop being resumed.
// RK_EVENT: *io describes the accepted step in which a terminal event fired (roots still to be found).
template <class TR>
__device__ __forceinline__ int rk_loop(const DevCfg& c, const EnvRegs<TR>& e, Integ<TR>& S, double (&g)[TR::NEV],
                                       bool rejected, EventIO<TR>* io, double (&pd)[TR::NP]) {
    const double T = c.T;
    for (;;) {
        const double min_step = min_step_at(S.t);
        if (S.h_abs < min_step && !rejected) S.h_abs = min_step;
        double Ax[7], Ay[7], xn, yn, thn, vxn, vyn, omn, cn, sn, err2, h, t_new, ipd[TR::NP];
        for (;;) {
            if (S.h_abs < min_step) return RK_FAILED;
            t_new = S.t + S.h_abs;
            if (t_new - T > 0) t_new = T;
            h = t_new - S.t;
            S.h_abs = fabs(h);
            rk_attempt<TR>(c, e, S, h, Ax, Ay, xn, yn, thn, vxn, vyn, omn, cn, sn, err2, pd, ipd);
            if (controller(err2, rejected, t_new < T, S.h_abs)) break;
        }
        rejected = false;
        double gn[TR::NEV];
        event_values_from_pd<TR>(c, xn, yn, omn, pd, gn);
        const uint32_t active = active_events<TR>(g, gn);
        if (active) {
            fill_event_io<TR>(*io, e, S, h, Ax, Ay, active);
            return RK_EVENT;
        }
#pragma unroll
        for (int i = 0; i < TR::NEV; ++i) g[i] = gn[i];
        S.x = xn; S.y = yn; S.th = thn; S.vx = vxn; S.vy = vyn; S.om = omn;
        S.cth = cn; S.sth = sn;
        S.Ax = Ax[6]; S.Ay = Ay[6];
        S.t = t_new;
        if (S.t - T >= 0) return RK_FINISHED;  // scipy:base.py:199
    }
}

// wrap_ship_angle (dynamic_model.py:179-180): numpy floor-mod by 2 pi; also flags non-finite states
__device__ __forceinline__ double wrap_angle(double th) {
    // th mod 2 pi with numpy's floor-mod semantics.  The heading enters a step in [0, 2 pi] and moves by
    // less than 0.5 rad, so one conditional add/subtract is the exact fmod result (Sterbenz); anything
    // else (states injected from outside) takes the general path.
    if (th >= -kTwoPi && th < 2 * kTwoPi) {
        if (th >= kTwoPi) th -= kTwoPi;
        else if (th < 0) th += kTwoPi;
        return th;
    }
    th = fmod(th, kTwoPi);
    if (th < 0) th += kTwoPi;
    return th;
}

template <class TR>
__device__ __forceinline__ void commit_state(EnvRegs<TR>& e, double x, double y, double th, double vx, double vy,
                                             double om, uint32_t& status) {
    th = wrap_angle(th);
    if (!(isfinite(x) && isfinite(y) && isfinite(th) && isfinite(vx) && isfinite(vy) && isfinite(om))) status |= 8u;
    e.x = x; e.y = y; e.th = th; e.vx = vx; e.vy = vy; e.om = om;
}

// ---- one env.step of the physics, everything inline (used by the K-fused kernel) ---------------------
// Returns done.  On exit e holds the new state (heading wrapped) and pd the planet distances there.
template <class TR>
__device__ __forceinline__ bool make_step(const DevCfg& c, EnvRegs<TR>& e, const Act& act, uint32_t& status,
                                          double (&pd)[TR::NP]) {
    Integ<TR> S;
    double g[TR::NEV];
    init_step<TR>(c, e, act, S, g, pd);
    EventIO<TR> io;
    const int rc = rk_loop<TR>(c, e, S, g, false, &io, pd);
    bool done = false;
    if (rc == RK_EVENT) {
        handle_events<TR>(c, &io);
        status |= io.status;
        commit_state<TR>(e, io.x, io.y, io.th, io.vx, io.vy, io.om, status);
        double gdummy[TR::NEV];
        event_values(c, e, e.x, e.y, e.om, gdummy, pd);
        done = true;
    } else {
        if (rc == RK_FAILED) status |= 1u;  // SG_ENV_STEP_TOO_SMALL
        commit_state<TR>(e, S.x, S.y, S.th, S.vx, S.vy, S.om, status);
    }
    return done;
}

// ---- observation (spaceship_env.py:113-140, kepler.py:172-187) ---------------------------------------
// Lidar = unit(obj - pos) * (|obj - pos| - r) * 2 / ws, written without atan2/cos/sin.
// observation when 1 / |planet - ship| is already known (from the event check at the same state)
template <class TR, typename OT>
__device__ __forceinline__ void make_obs_inv(const DevCfg& c, const EnvRegs<TR>& e, double cs, double sn,
                                             const double (&ipd)[TR::NP], OT* row) {
    row[0] = (OT)e.x;
    row[1] = (OT)e.y;
    row[2] = (OT)cs;
    row[3] = (OT)sn;
    row[4] = (OT)e.vx;
    row[5] = (OT)e.vy;
    row[6] = (OT)e.om;
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) {
            const double k = (1.0 - c.r0 * ipd[p]) * c.two_over_ws;
            row[7 + 2 * p] = (OT)((e.px[p] - e.x) * k);
            row[8 + 2 * p] = (OT)((e.py[p] - e.y) * k);
        }
        row[7 + 2 * TR::NP] = (OT)((e.gx - e.x) * c.two_over_ws);
        row[8 + 2 * TR::NP] = (OT)((e.gy - e.y) * c.two_over_ws);
    } else if (TR::TASK == TASK_KEPLER) {
        row[7] = (OT)e.ang;
        row[8] = (OT)e.ecc;
        row[9] = (OT)c.orbit_a;
    }
}

template <class TR, typename OT>
__device__ __forceinline__ void make_obs_cs(const DevCfg& c, const EnvRegs<TR>& e, double cs, double sn, OT* row) {
    row[0] = (OT)e.x;
    row[1] = (OT)e.y;
    row[2] = (OT)cs;
    row[3] = (OT)sn;
    row[4] = (OT)e.vx;
    row[5] = (OT)e.vy;
    row[6] = (OT)e.om;
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) {
            double dx = e.px[p] - e.x, dy = e.py[p] - e.y;
            double d2 = dx * dx + dy * dy;
            double k = (1.0 - c.r0 * fast_rsqrt(d2)) * c.two_over_ws;
            row[7 + 2 * p] = (OT)(dx * k);
            row[8 + 2 * p] = (OT)(dy * k);
        }
        row[7 + 2 * TR::NP] = (OT)((e.gx - e.x) * c.two_over_ws);
        row[8 + 2 * TR::NP] = (OT)((e.gy - e.y) * c.two_over_ws);
    } else if (TR::TASK == TASK_KEPLER) {
        row[7] = (OT)e.ang;
        row[8] = (OT)e.ecc;
        row[9] = (OT)c.orbit_a;
    }
}

template <class TR, typename OT>
__device__ __forceinline__ void make_obs(const DevCfg& c, const EnvRegs<TR>& e, OT* row) {
    double sn, cs;
    sincos(e.th, &sn, &cs);
    make_obs_cs<TR, OT>(c, e, cs, sn, row);
}

// ---- rewards ------------------------------------------------------------------------------------------
// GoalEnv._reward without the sparse term (goal.py:147-165,204-227); cur = |goal - pos|.
template <class TR>
__device__ __forceinline__ double goal_reward(const DevCfg& c, const EnvRegs<TR>& e, double lx, double ly,
                                              const double (&pd)[TR::NP], double& cur) {
    double dx = e.gx - e.x, dy = e.gy - e.y;
    cur = fast_sqrt0(dx * dx + dy * dy);
    double ldx = e.gx - lx, ldy = e.gy - ly;
    double last = fast_sqrt0(ldx * ldx + ldy * ldy);
    double gv = (last - cur) * 100.0;
    double mind = INFINITY, cx = 0, cy = 0;
#pragma unroll
    for (int p = 0; p < TR::NP; ++p) {
        bool closer = pd[p] < mind;
        mind = closer ? pd[p] : mind;
        cx = closer ? e.px[p] : cx;
        cy = closer ? e.py[p] : cy;
    }
    double sum_safety = 0.0;
    if (mind - c.r0 < c.danger_zone) {
        double qx = lx - cx, qy = ly - cy;
        double prev = fast_sqrt0(qx * qx + qy * qy);
        if (prev > mind) sum_safety = -100.0 * (prev - mind);
    }
    return c.survival + c.goal_vel * gv + c.safety * sum_safety;
}

// KeplerEnv._dense_reward5 (kepler.py:43-150); atan2/cos replaced by the equivalent ratios.
template <class TR>
__device__ __forceinline__ double kepler_reward(const DevCfg& c, const EnvRegs<TR>& e, const Act& act) {
    const double a = c.orbit_a, ecc = e.ecc;
    // no FMA contraction here: for a circular orbit a*a - b*b must cancel to exactly 0 as it does in
    // the reference (a fused a*a - b*b leaves the rounding error of b*b, and its sqrt is ~3e-9)
    const double aa = __dmul_rn(a, a);
    double b = sqrt(__dmul_rn(aa, __dsub_rn(1.0, __dmul_rn(ecc, ecc))));
    double cc = sqrt(__dsub_rn(aa, __dmul_rn(b, b)));
    double sa, ca;
    sincos(e.ang, &sa, &ca);
    double w0 = ca * e.x + sa * e.y - cc;  // _rotate, then shift by the focal distance
    double w1 = -sa * e.x + ca * e.y;
    double nw2 = w0 * w0 + w1 * w1;
    double inw = fast_rsqrt(nw2);
    double cur_rad = nw2 * inw;
    double ecos = ecc * (w0 * inw);  // cos(atan2(w1, w0))
    double target_rad = b * fast_rsqrt(1 - ecos * ecos);
    double p0 = w0 * target_rad * inw, p1 = w1 * target_rad * inw;  // projection onto the orbit
    const double ab = fast_div(a, b);
    double v0 = -ab * p1, v1 = fast_div(b, a) * p0;
    double rx = p0 + cc;
    double r = fast_sqrt(rx * rx + p1 * p1);
    double ov = fast_sqrt0(c.gm0 * (2 * fast_rcp(r) - fast_rcp(a)));  // vis-viva
    double sc = ov * fast_rsqrt(v0 * v0 + v1 * v1);
    v0 *= sc;
    v1 *= sc;
    double t0 = ca * v0 - sa * v1;  // _rotate by -angle
    double t1 = sa * v0 + ca * v1;
    double den = c.radC * fabs(cur_rad - target_rad) + fabs(t0 - e.vx) + fabs(t1 - e.vy) + act.act_term + c.numC;
    return fast_div(c.numC, den);
}

// ---- hexagonal tiling state (hexagonal_tiling.py) ---------------------------------------------------------
struct Tiles {
    uint32_t case_b, flip_xy, ship_tile, goal_tile, n_free;  // goal_tile 63 = None
    unsigned long long fw0, fw1, fw2, fw3;                   // free multiset, 4 bits per entry (registers, no indexing)
    double cs0, cs1, cs2, cs3;                               // column shifts
    __device__ unsigned long long word(uint32_t w) const { return w == 0 ? fw0 : (w == 1 ? fw1 : (w == 2 ? fw2 : fw3)); }
    __device__ uint32_t get(uint32_t k) const { return (uint32_t)(word(k >> 4) >> ((k & 15u) * 4u)) & 15u; }
    __device__ void set(uint32_t k, uint32_t v) {
        const uint32_t w = k >> 4, sh = (k & 15u) * 4u;
        const unsigned long long m = 15ull << sh;
        const unsigned long long nw = (word(w) & ~m) | ((unsigned long long)v << sh);
        fw0 = (w == 0) ? nw : fw0;
        fw1 = (w == 1) ? nw : fw1;
        fw2 = (w == 2) ? nw : fw2;
        fw3 = (w == 3) ? nw : fw3;
    }
    __device__ double shift(int col) const { return col == 0 ? cs0 : (col == 1 ? cs1 : (col == 2 ? cs2 : cs3)); }
    __device__ uint32_t pack() const {
        return case_b | (flip_xy << 1) | (ship_tile << 2) | (goal_tile << 8) | (n_free << 14);
    }
    __device__ void unpack(uint32_t w) {
        case_b = w & 1u; flip_xy = (w >> 1) & 1u; ship_tile = (w >> 2) & 63u; goal_tile = (w >> 8) & 63u;
        n_free = (w >> 14) & 127u;
    }
};

__device__ inline void load_tiles(const DevCfg& c, long long i, Tiles& t) {
    t.unpack(c.tile_word[i]);
    t.fw0 = c.free_tiles[i];
    t.fw1 = c.free_tiles[c.stride + i];
    t.fw2 = c.free_tiles[2 * c.stride + i];
    t.fw3 = c.free_tiles[3 * c.stride + i];
    t.cs0 = c.col_shift[i];
    t.cs1 = c.cols > 1 ? c.col_shift[c.stride + i] : 0.0;
    t.cs2 = c.cols > 2 ? c.col_shift[2 * c.stride + i] : 0.0;
    t.cs3 = c.cols > 3 ? c.col_shift[3 * c.stride + i] : 0.0;
}
__device__ inline void store_tiles(const DevCfg& c, long long i, const Tiles& t, bool with_shift) {
    c.tile_word[i] = t.pack();
    c.free_tiles[i] = t.fw0;
    c.free_tiles[c.stride + i] = t.fw1;
    c.free_tiles[2 * c.stride + i] = t.fw2;
    c.free_tiles[3 * c.stride + i] = t.fw3;
    if (with_shift) {
        c.col_shift[i] = t.cs0;
        if (c.cols > 1) c.col_shift[c.stride + i] = t.cs1;
        if (c.cols > 2) c.col_shift[2 * c.stride + i] = t.cs2;
        if (c.cols > 3) c.col_shift[3 * c.stride + i] = t.cs3;
    }
}

// HexagonalTiling._tile_center_pos (hexagonal_tiling.py:136-158)
static __device__ __noinline__ void tile_center(const DevCfg& c, const Tiles& t, int tile, double& ox, double& oy) {
    int row = tile / c.cols, col = tile % c.cols;
    double zero_x = -c.ws / 2 + c.hex_width / 2;
    double zero_y = c.ws / 2 - c.hex_height / 2;
    if (t.case_b) zero_y -= c.hex_height / 2;
    double x_shift = col * 1.5 * c.hex_a + t.shift(col);
    double y_rows = -row * c.hex_height;
    double y_cols = -(col % 2) * c.hex_height / 2;
    if (t.case_b) y_cols *= -1;
    double cx = zero_x + x_shift, cy = zero_y + (y_rows + y_cols);
    ox = t.flip_xy ? cy : cx;
    oy = t.flip_xy ? cx : cy;
}

// choice(n, size=k, replace=False) as a partial Fisher-Yates (uniform ordered k-subset).  For n <= 16 the
// permutation lives in one 64-bit register, 4 bits per element; longer free lists (rare) use a local array.
template <int KMAX>
__device__ inline void choice_no_replace(Rng& rng, int n, int k, int (&out)[KMAX]) {
    if (n <= 16) {
        unsigned long long perm = 0xFEDCBA9876543210ull;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
            if (i < k) {
                int j = i + (int)(rng.uniform() * (n - i));
                if (j > n - 1) j = n - 1;
                const unsigned long long ai = (perm >> (4 * i)) & 15ull, aj = (perm >> (4 * j)) & 15ull;
                const unsigned long long dx = ai ^ aj;
                perm ^= (dx << (4 * i)) | (dx << (4 * j));  // swap the two nibbles (no-op when i == j)
                out[i] = (int)aj;
            }
        }
        return;
    }
    int arr[64];
    for (int i = 0; i < n; ++i) arr[i] = i;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        if (i < k) {
            int j = i + (int)(rng.uniform() * (n - i));
            if (j > n - 1) j = n - 1;
            int tmp = arr[i];
            arr[i] = arr[j];
            arr[j] = tmp;
            out[i] = arr[i];
        }
    }
}

// GoalEnv._sample_position_outside_one_planet (goal.py:90-99) with helpers.get_max_dist_in_direction (helpers.py:38-46):
// a point on a uniformly drawn ray from the planet, between planet_radius + clearance and the world box shrunk by
// the clearance.  IEEE divisions as in NumPy (a zero direction component gives +-inf / nan candidates, which the
// `> 0` filter treats as the reference's does).  Works on either draw session (Rng / Draws).
template <class RNG>
__device__ inline void sample_outside_one_planet(const DevCfg& c, RNG& rng, double px, double py, double clearance,
                                                 double& ox, double& oy, uint32_t& status) {
    const double max_obj = c.half_ws - clearance;
    const double ang = kTwoPi * rng.uniform();
    double sn, cs;
    sincos_cold(ang, &sn, &cs);
    const double cand[4] = {(max_obj - px) / cs, (-max_obj - px) / cs, (max_obj - py) / sn, (-max_obj - py) / sn};
    double dmax = INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cand[k] > 0 && cand[k] < dmax) dmax = cand[k];
    const double dmin = c.r0 + clearance;
    if (!(dmin < dmax)) status |= 16u;  // SG_ENV_PLACEMENT: the reference asserts (goal.py:96)
    const double dist = dmin + (dmax - dmin) * rng.uniform();
    ox = px + cs * dist;
    oy = py + sn * dist;
}

// env.reset of the one-planet GoalEnv (goal.py:101-109,133-145 with the unpacking of :134 fixed: the planet takes
// the sampled planet position, the first goal sample is drawn and dropped, `_resample_goal` draws the goal).
// out: x, y, th, vx, vy, om | planet (x, y) | goal (x, y)
template <class RNG>
__device__ inline void reset_goal_one_planet(const DevCfg& c, RNG& rng, double* out, uint32_t& status) {
    const double max_pos = c.half_ws - c.r0;
    const double clear = fmax(c.ship_radius, c.goal_radius);
    const double pdist = (max_pos - 2 * clear) * rng.uniform();
    const double pang = kTwoPi * rng.uniform();
    double sn, cs;
    sincos_cold(pang, &sn, &cs);
    const double px = cs * pdist, py = sn * pdist;
    out[6] = px;
    out[7] = py;
    double ux, uy;
    sample_outside_one_planet(c, rng, px, py, c.ship_radius, out[0], out[1], status);
    sample_outside_one_planet(c, rng, px, py, c.goal_radius, ux, uy, status);
    sample_outside_one_planet(c, rng, px, py, c.goal_radius, out[8], out[9], status);
    out[2] = kTwoPi * rng.uniform();
    out[3] = rng.normal() * 0.07;
    out[4] = rng.normal() * 0.07;
    const double maxw = 0.7 * c.om_max;
    const double w = rng.normal() * maxw / 3;
    out[5] = fmin(fmax(w, -maxw), maxw);
}

// HexagonalTiling.find_new_goal (hexagonal_tiling.py:95-134)
static __device__ __noinline__ void find_new_goal(const DevCfg& c, Tiles& t, Rng& rng, double& gx, double& gy, uint32_t& status) {
    if (t.goal_tile != 63u) {
        if (t.n_free < 64u) {
            t.set(t.n_free, t.ship_tile);
            t.n_free += 1;
        } else {
            status |= 4u;  // SG_ENV_FREE_OVERFLOW
        }
        t.ship_tile = t.goal_tile;
    }
    if (rng.uniform() < 0.25) {
        t.goal_tile = t.ship_tile;
    } else {
        int k = t.n_free < 3u ? (int)t.n_free : 3;
        int cand[3];
        choice_no_replace<3>(rng, (int)t.n_free, k, cand);
        int best = -1, best_idx = 0;
        int srow = (int)t.ship_tile / c.cols, scol = (int)t.ship_tile % c.cols;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < k) {
                int tnr = (int)t.get((uint32_t)cand[i]);
                int taxi = abs(tnr / c.cols - srow) + abs(tnr % c.cols - scol);
                if (taxi > best) {
                    best = taxi;
                    best_idx = cand[i];
                }
            }
        }
        t.goal_tile = t.get((uint32_t)best_idx);
        for (uint32_t i = (uint32_t)best_idx; i + 1 < t.n_free; ++i) t.set(i, t.get(i + 1));
        t.n_free -= 1;
        t.set(t.n_free, 0u);  // keep vacated entries zero: the packed words are then canonical
    }
    double cx, cy;
    tile_center(c, t, (int)t.goal_tile, cx, cy);
    double R = c.hex_height / 2 - c.goal_radius;
    double ang = kTwoPi * rng.uniform();
    double rr = sqrt(rng.uniform() * (R * R));
    double sn, cs;
    sincos_cold(ang, &sn, &cs);
    gx = cx + rr * cs;
    gy = cy + rr * sn;
}

// Goal resample on a reached goal (goal.py:155-157): cold, touches the tile state in HBM.
template <class TR>
static __device__ __noinline__ void resample_goal(const DevCfg& c, long long i, Rng* rng, double* gxy, uint32_t* status) {
    if (!TR::TILED) {  // one planet: goal.py:122-123
        sample_outside_one_planet(c, *rng, c.planets[i], c.planets[c.stride + i], c.goal_radius, gxy[0], gxy[1], *status);
        return;
    }
    Tiles t;
    load_tiles(c, i, t);
    find_new_goal(c, t, *rng, gxy[0], gxy[1], *status);
    store_tiles(c, i, t, false);
}

// env.reset (goal.py:133-145 + hexagonal_tiling.py:53-93 | kepler.py:233-267).
// out: x, y, th, vx, vy, om; Goal: planets (x0, y0, ...) at [6..], goal at [6+2P]; Kepler: (ecc, ang) at [6..7]
// when randomize (left untouched otherwise).
template <class TR>
static __device__ __noinline__ void reset_env(const DevCfg& c, long long i, Rng* rngp, double* out, uint32_t* status) {
    Rng& rng = *rngp;
    if (TR::TASK == TASK_GOAL && !TR::TILED) {
        reset_goal_one_planet(c, rng, out, *status);
    } else if (TR::TASK == TASK_GOAL) {
        constexpr int P = TR::NP;
        const int n_tiles = c.rows * c.cols;
        Tiles t;
        t.fw0 = t.fw1 = t.fw2 = t.fw3 = 0;
        t.goal_tile = 63u;
        t.case_b = rng.uniform() < 0.5;
        t.flip_xy = rng.uniform() < 0.5;
        double acc = rng.uniform();  // cumsum of `cols` uniforms (cols >= 2), rescaled so the last one is free_x
        t.cs0 = acc;
        acc += rng.uniform();
        t.cs1 = acc;
        t.cs2 = t.cs3 = 0.0;
        if (c.cols > 2) {
            acc += rng.uniform();
            t.cs2 = acc;
        }
        if (c.cols > 3) {
            acc += rng.uniform();
            t.cs3 = acc;
        }
        const double sc = c.free_x / acc;
        t.cs0 *= sc; t.cs1 *= sc; t.cs2 *= sc; t.cs3 *= sc;
        int tiles[P + 1];
        bool diag = false;
        if (P == 2) diag = rng.uniform() < 0.25;
        if (diag) {
            const int cases[4][3] = {{1, 0, 3}, {2, 0, 3}, {0, 1, 2}, {3, 1, 2}};
            int k = (int)(rng.uniform() * 4);
            if (k > 3) k = 3;
            for (int j = 0; j < 3 && j < P + 1; ++j) tiles[j] = cases[k][j];
        } else {
            choice_no_replace<P + 1>(rng, n_tiles, P + 1, tiles);
        }
        t.ship_tile = (uint32_t)tiles[0];
        t.n_free = 0;
        for (int k = 0; k < n_tiles; ++k) {
            bool used = false;
            for (int j = 0; j < P + 1; ++j) used |= (tiles[j] == k);
            if (!used) {
                t.set(t.n_free, (uint32_t)k);
                t.n_free += 1;
            }
        }
        double ang[P + 1];
        for (int j = 0; j < P + 1; ++j) ang[j] = kTwoPi * rng.uniform();
        for (int j = 0; j < P + 1; ++j) {
            double radius = (j == 0) ? c.ship_radius : c.r0;
            double R = c.hex_height / 2 - radius;
            double rr = sqrt(rng.uniform() * (R * R));
            double cx, cy, sn, cs;
            tile_center(c, t, tiles[j], cx, cy);
            sincos_cold(ang[j], &sn, &cs);
            double ox = cx + rr * cs, oy = cy + rr * sn;
            if (j == 0) {
                out[0] = ox;
                out[1] = oy;
            } else {
                out[6 + 2 * (j - 1)] = ox;
                out[7 + 2 * (j - 1)] = oy;
            }
        }
        find_new_goal(c, t, rng, out[6 + 2 * P], out[7 + 2 * P], *status);
        store_tiles(c, i, t, true);
        out[2] = kTwoPi * rng.uniform();
        out[3] = rng.normal() * 0.07;
        out[4] = rng.normal() * 0.07;
        double maxw = 0.7 * c.om_max;
        double w = rng.normal() * maxw / 3;
        out[5] = fmin(fmax(w, -maxw), maxw);
    } else {
        double pang = kTwoPi * rng.uniform();
        double dist = c.kep_lo + (c.kep_hi - c.kep_lo) * rng.uniform();
        double sn, cs;
        sincos_cold(pang, &sn, &cs);
        out[0] = cs * dist;
        out[1] = sn * dist;
        out[2] = kTwoPi * rng.uniform();
        if (TR::TASK == TASK_KEPLER && c.randomize) {
            out[6] = rng.uniform() * 0.7;
            out[7] = rng.uniform() * 2 * kPi;
        }
        out[3] = rng.normal() * c.reset_v_std;
        out[4] = rng.normal() * c.reset_v_std;
        double maxw = 0.7 * c.om_max;
        double w = rng.normal() * maxw / c.reset_om_div;
        out[5] = fmin(fmax(w, -maxw), maxw);
    }
}

}  // namespace sg
