// sg_tail.cuh — the rare paths of an env step, written for LATENCY: terminal-event roots on the dense output
// (scipy:ivp.py:52-77,679-699; scipy:rk.py:715-737), goal resampling (goal.py:155-157, hexagonal_tiling.py:95-134)
// and env.reset (goal.py:133-145, hexagonal_tiling.py:53-93, kepler.py:233-267, do_not_crash.py:34-45).
//
// These run for ~1.5 % of the envs of a step, in a kernel that sits at the end of the step's dependency chain, so
// what counts is the length of one lane's serial instruction chain, not throughput:
//   * the env's Philox stream is counter-based, so every block a reset can consume is computed up front, all
//     blocks independent (instruction-level parallelism instead of ~16 dependent 10-round chains behind calls),
//     into shared memory; the draw session then reads uniforms by position;
//   * the roots of the event functions on the quartic dense output are found by a bracketed Newton iteration
//     with the analytic derivative of the interpolant (3-5 evaluations) instead of Brent's method (8-12
//     evaluations plus its division-heavy bookkeeping) — SciPy's brentq is asked for xtol = rtol = 4 eps, and
//     both iterations converge to the same root of the same polynomial to that tolerance;
//   * sin / cos of the reset angles (all in [0, 2 pi)) by quadrant reduction + Taylor polynomials, no slow path;
//   * the tiling shape (rows = cols = n_planets for the GoalEnv family) is a compile-time constant.
#pragma once

#include "sg_device.cuh"

namespace sg {

// sincos_2pi (sin / cos for angles in [0, 2 pi]: reset angles, Box-Muller angles) lives in sg_device.cuh.

// ---- draw session on precomputed Philox blocks ------------------------------------------------------------------
// Same stream as Rng (sg_device.cuh): block b of env stream 0 = Philox(counter (env_lo, env_hi, ctr + b, 0), key seed)
// gives the doubles 2b, 2b+1.  The buffer of lane l holds element k at u[k * 32] (conflict-free for a warp).
constexpr int kDrawBlocks = 16;  // buffer size: 32 doubles, the longest session (a 4-planet reset: exactly 32 draws)

// Philox blocks the longest draw session of an env family can consume.  GoalEnv reset (Appendix C of SURVEY.md):
// case_b, flip_xy | `cols` column shifts | 2 planets: diagonal test + (case index | 3 tiles), else P+1 tiles |
// P+1 angles | P+1 radii | find_new_goal: 1 + 3 + 2 | heading | 3 normals = 2 Box-Muller pairs.  Kepler /
// DoNotCrash: 3 (+2 when randomize) + 4.  A goal resample alone draws at most 6.
template <class TR>
__host__ __device__ constexpr int draw_blocks() {
    return TR::TASK == TASK_GOAL ? (2 + TR::NP + (TR::NP == 2 ? 4 : TR::NP + 1) + 2 * (TR::NP + 1) + 6 + 1 + 4 + 1) / 2 : 5;
}

template <int NB>
__device__ __forceinline__ void philox_fill(unsigned long long seed, unsigned long long env, uint32_t ctr, double* u) {
    static_assert(NB <= kDrawBlocks, "draw buffer too small");
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), e0 = (uint32_t)env, e1 = (uint32_t)(env >> 32);
#pragma unroll 4
    for (int b = 0; b < NB; ++b) {
        uint32_t o[4];
        philox4x32_10(e0, e1, ctr + (uint32_t)b, 0u, k0, k1, o);
        const unsigned long long lo = ((unsigned long long)o[1] << 32) | o[0];
        const unsigned long long hi = ((unsigned long long)o[3] << 32) | o[2];
        u[(2 * b) * 32] = (double)(lo >> 11) * (1.0 / 9007199254740992.0);
        u[(2 * b + 1) * 32] = (double)(hi >> 11) * (1.0 / 9007199254740992.0);
    }
}

struct Draws {
    const double* u;
    int pos;
    bool have_n;
    double nbuf;
    __device__ __forceinline__ void open(const double* buf) {
        u = buf;
        pos = 0;
        have_n = false;
        nbuf = 0.0;
    }
    __device__ __forceinline__ double uniform() {
        const int p = pos < 2 * kDrawBlocks - 1 ? pos : 2 * kDrawBlocks - 1;
        pos += 1;
        return u[p * 32];
    }
    __device__ __forceinline__ double normal() {  // Box-Muller, second value cached (as Rng::normal)
        if (have_n) {
            have_n = false;
            return nbuf;
        }
        const double u1 = 1.0 - uniform();
        const double u2 = uniform();
        const double r = fast_sqrt0(-2.0 * log(u1));
        double s, c;
        sincos_2pi(kTwoPi * u2, s, c);
        nbuf = r * s;
        have_n = true;
        return r * c;
    }
    __device__ __forceinline__ uint32_t blocks_used() const { return (uint32_t)(pos + 1) >> 1; }
};

// ---- hexagonal tiling with compile-time shape --------------------------------------------------------------------
// HexagonalTiling._tile_center_pos (hexagonal_tiling.py:136-158); COLS = the tiling's column count
template <int COLS>
__device__ __forceinline__ void tile_center_t(const DevCfg& c, const Tiles& t, int tile, double& ox, double& oy) {
    const int row = tile / COLS, col = tile % COLS;
    const double zero_x = -c.ws / 2 + c.hex_width / 2;
    double zero_y = c.ws / 2 - c.hex_height / 2;
    if (t.case_b) zero_y -= c.hex_height / 2;
    const double x_shift = col * 1.5 * c.hex_a + t.shift(col);
    const double y_rows = -row * c.hex_height;
    double y_cols = -(col % 2) * c.hex_height / 2;
    if (t.case_b) y_cols *= -1;
    const double cx = zero_x + x_shift, cy = zero_y + (y_rows + y_cols);
    ox = t.flip_xy ? cy : cx;
    oy = t.flip_xy ? cx : cy;
}

// choice(n, size=k, replace=False), n <= 16: the permutation is one 64-bit register (as choice_no_replace)
template <int KMAX>
__device__ __forceinline__ void choice16(Draws& rng, int n, int k, int (&out)[KMAX]) {
    unsigned long long perm = 0xFEDCBA9876543210ull;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        if (i < k) {
            int j = i + (int)(rng.uniform() * (n - i));
            if (j > n - 1) j = n - 1;
            const unsigned long long ai = (perm >> (4 * i)) & 15ull, aj = (perm >> (4 * j)) & 15ull;
            const unsigned long long dx = ai ^ aj;
            perm ^= (dx << (4 * i)) | (dx << (4 * j));
            out[i] = (int)aj;
        }
    }
}

// Candidate pick of find_new_goal for a free multiset of more than 16 entries (the list grows by one whenever a new
// goal lands on the ship's own tile, hexagonal_tiling.py:101-108: only long episodes of a goal-seeking policy get
// here).  Same draws as the short form; the permutation lives in a local array.
static __device__ __noinline__ void pick_goal_long(int cols, Tiles& t, Draws& rng) {
    const int n = (int)t.n_free;
    int arr[64], cand[3];
    for (int i = 0; i < n; ++i) arr[i] = i;
    for (int i = 0; i < 3; ++i) {
        int j = i + (int)(rng.uniform() * (n - i));
        if (j > n - 1) j = n - 1;
        const int tmp = arr[i];
        arr[i] = arr[j];
        arr[j] = tmp;
        cand[i] = arr[i];
    }
    int best = -1, best_idx = 0;
    const int srow = (int)t.ship_tile / cols, scol = (int)t.ship_tile % cols;
    for (int i = 0; i < 3; ++i) {
        const int tnr = (int)t.get((uint32_t)cand[i]);
        const int taxi = abs(tnr / cols - srow) + abs(tnr % cols - scol);
        if (taxi > best) {
            best = taxi;
            best_idx = cand[i];
        }
    }
    t.goal_tile = t.get((uint32_t)best_idx);
    for (uint32_t i = (uint32_t)best_idx; i + 1 < t.n_free; ++i) t.set(i, t.get(i + 1));
    t.n_free -= 1;
    t.set(t.n_free, 0u);  // vacated entries stay zero: the packed words are canonical
}

// HexagonalTiling.find_new_goal (hexagonal_tiling.py:95-134)
template <int COLS>
__device__ __forceinline__ void find_new_goal_t(const DevCfg& c, Tiles& t, Draws& rng, double& gx, double& gy,
                                                uint32_t& status) {
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
    } else if (t.n_free <= 16u) {
        const int k = t.n_free < 3u ? (int)t.n_free : 3;
        int cand[3];
        choice16<3>(rng, (int)t.n_free, k, cand);
        int best = -1, best_idx = 0;
        const int srow = (int)t.ship_tile / COLS, scol = (int)t.ship_tile % COLS;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < k) {
                const int tnr = (int)t.get((uint32_t)cand[i]);
                const int taxi = abs(tnr / COLS - srow) + abs(tnr % COLS - scol);
                if (taxi > best) {
                    best = taxi;
                    best_idx = cand[i];
                }
            }
        }
        t.goal_tile = t.get((uint32_t)best_idx);
        // remove entry best_idx: the entries above it move down one nibble (the list fits word 0: n_free <= 16)
        const unsigned long long keep = (best_idx == 0) ? 0ull : (~0ull >> (64 - 4 * best_idx));
        t.fw0 = (t.fw0 & keep) | ((t.fw0 >> 4) & ~keep);
        t.n_free -= 1;
    } else {
        pick_goal_long(COLS, t, rng);
    }
    double cx, cy;
    tile_center_t<COLS>(c, t, (int)t.goal_tile, cx, cy);
    const double R = c.hex_height / 2 - c.goal_radius;
    const double ang = kTwoPi * rng.uniform();
    const double rr = fast_sqrt0(rng.uniform() * (R * R));
    double sn, cs;
    sincos_2pi(ang, sn, cs);
    gx = cx + rr * cs;
    gy = cy + rr * sn;
}

// env.reset on a draw session.  out: x, y, th, vx, vy, om; Goal: planets (x0, y0, ...) at [6..], goal at [6+2P];
// Kepler: (ecc, ang) at [6..7] when randomize.  Goal: the tile state is returned in t (the caller stores it).
template <class TR>
__device__ __forceinline__ void reset_env_t(const DevCfg& c, Draws& rng, double (&out)[8 + 2 * TR::NP], Tiles& t,
                                            uint32_t& status) {
    if (TR::TASK == TASK_GOAL && !TR::TILED) {
        reset_goal_one_planet(c, rng, out, status);
    } else if (TR::TASK == TASK_GOAL) {
        constexpr int P = TR::NP, COLS = TR::NP, n_tiles = TR::NP * TR::NP;
        t.fw0 = t.fw1 = t.fw2 = t.fw3 = 0;
        t.goal_tile = 63u;
        t.case_b = rng.uniform() < 0.5;
        t.flip_xy = rng.uniform() < 0.5;
        double acc = rng.uniform();  // cumsum of `cols` uniforms, rescaled so that the last one is free_x
        t.cs0 = acc;
        acc += rng.uniform();
        t.cs1 = acc;
        t.cs2 = t.cs3 = 0.0;
        if (COLS > 2) {
            acc += rng.uniform();
            t.cs2 = acc;
        }
        if (COLS > 3) {
            acc += rng.uniform();
            t.cs3 = acc;
        }
        const double sc = c.free_x / acc;
        t.cs0 *= sc; t.cs1 *= sc; t.cs2 *= sc; t.cs3 *= sc;
        int tiles[P + 1];
        bool diag = false;
        if (P == 2) diag = rng.uniform() < 0.25;
        if (diag) {
            int k = (int)(rng.uniform() * 4);
            if (k > 3) k = 3;
            // cases {1,0,3} {2,0,3} {0,1,2} {3,1,2} (hexagonal_tiling.py:72-79), one nibble per tile
            const uint32_t packed = k == 0 ? 0x301u : (k == 1 ? 0x302u : (k == 2 ? 0x210u : 0x213u));
#pragma unroll
            for (int j = 0; j < 3 && j < P + 1; ++j) tiles[j] = (int)((packed >> (4 * j)) & 15u);
        } else {
            choice16<P + 1>(rng, n_tiles, P + 1, tiles);
        }
        t.ship_tile = (uint32_t)tiles[0];
        uint32_t used = 0;
#pragma unroll
        for (int j = 0; j < P + 1; ++j) used |= 1u << tiles[j];
        t.n_free = 0;
#pragma unroll
        for (int k = 0; k < n_tiles; ++k) {
            if (!((used >> k) & 1u)) {
                t.fw0 |= (unsigned long long)k << (4 * t.n_free);
                t.n_free += 1;
            }
        }
        double ang[P + 1];
#pragma unroll
        for (int j = 0; j < P + 1; ++j) ang[j] = kTwoPi * rng.uniform();
#pragma unroll
        for (int j = 0; j < P + 1; ++j) {
            const double radius = (j == 0) ? c.ship_radius : c.r0;
            const double R = c.hex_height / 2 - radius;
            const double rr = fast_sqrt0(rng.uniform() * (R * R));
            double cx, cy, sn, cs;
            tile_center_t<COLS>(c, t, tiles[j], cx, cy);
            sincos_2pi(ang[j], sn, cs);
            const double ox = cx + rr * cs, oy = cy + rr * sn;
            if (j == 0) {
                out[0] = ox;
                out[1] = oy;
            } else {
                out[6 + 2 * (j - 1)] = ox;
                out[7 + 2 * (j - 1)] = oy;
            }
        }
        find_new_goal_t<COLS>(c, t, rng, out[6 + 2 * P], out[7 + 2 * P], status);
        out[2] = kTwoPi * rng.uniform();
        out[3] = rng.normal() * 0.07;
        out[4] = rng.normal() * 0.07;
        const double maxw = 0.7 * c.om_max;
        const double w = rng.normal() * maxw / 3;
        out[5] = fmin(fmax(w, -maxw), maxw);
    } else {
        const double pang = kTwoPi * rng.uniform();
        const double dist = c.kep_lo + (c.kep_hi - c.kep_lo) * rng.uniform();
        double sn, cs;
        sincos_2pi(pang, sn, cs);
        out[0] = cs * dist;
        out[1] = sn * dist;
        out[2] = kTwoPi * rng.uniform();
        if (TR::TASK == TASK_KEPLER && c.randomize) {
            out[6] = rng.uniform() * 0.7;
            out[7] = rng.uniform() * 2 * kPi;
        }
        out[3] = rng.normal() * c.reset_v_std;
        out[4] = rng.normal() * c.reset_v_std;
        const double maxw = 0.7 * c.om_max;
        const double w = rng.normal() * maxw / c.reset_om_div;
        out[5] = fmin(fmax(w, -maxw), maxw);
    }
}

// ---- terminal events on the dense output ---------------------------------------------------------------------------
// Interpolant of an accepted step (scipy:rk.py:178-180,715-737): y(t) = y_old + h Q [s, s^2, s^3, s^4], s = (t - t_old) / h.
// The heading and its rate are exact polynomials in time (sg_device.cuh), so Q is only kept for x, y, vx, vy.
template <class TR>
struct DenseQ {
    double x0, y0, th0, vx0, vy0, om0, alpha, h, t_old;
    double Qx[4], Qy[4], Qvx[4], Qvy[4];
};

template <class TR>
__device__ __forceinline__ void dense_build(DenseQ<TR>& d, const double (&Ax)[7], const double (&Ay)[7]) {
    const double h = d.h, vx = d.vx0, vy = d.vy0;
    // stage velocities (rows 0-1 of K): V_i = v + h sum_j a_ij A_j
    double Vx[7], Vy[7];
    Vx[0] = vx; Vy[0] = vy;
    Vx[1] = vx + h * (a21 * Ax[0]);
    Vy[1] = vy + h * (a21 * Ay[0]);
    Vx[2] = vx + h * (a31 * Ax[0] + a32 * Ax[1]);
    Vy[2] = vy + h * (a31 * Ay[0] + a32 * Ay[1]);
    Vx[3] = vx + h * (a41 * Ax[0] + a42 * Ax[1] + a43 * Ax[2]);
    Vy[3] = vy + h * (a41 * Ay[0] + a42 * Ay[1] + a43 * Ay[2]);
    Vx[4] = vx + h * (a51 * Ax[0] + a52 * Ax[1] + a53 * Ax[2] + a54 * Ax[3]);
    Vy[4] = vy + h * (a51 * Ay[0] + a52 * Ay[1] + a53 * Ay[2] + a54 * Ay[3]);
    Vx[5] = vx + h * (a61 * Ax[0] + a62 * Ax[1] + a63 * Ax[2] + a64 * Ax[3] + a65 * Ax[4]);
    Vy[5] = vy + h * (a61 * Ay[0] + a62 * Ay[1] + a63 * Ay[2] + a64 * Ay[3] + a65 * Ay[4]);
    Vx[6] = vx + h * (b1 * Ax[0] + b3 * Ax[2] + b4 * Ax[3] + b5 * Ax[4] + b6 * Ax[5]);
    Vy[6] = vy + h * (b1 * Ay[0] + b3 * Ay[2] + b4 * Ay[3] + b5 * Ay[4] + b6 * Ay[5]);
    // Q = K^T P with the zero entries of P (row 2, column 1 below row 1) left out; same summation order as
    // the dot product SciPy does (k ascending)
#define SG_QCOL(V, j, Pa, Pc, Pd, Pe, Pf, Pg) (V[0] * (Pa) + V[2] * (Pc) + V[3] * (Pd) + V[4] * (Pe) + V[5] * (Pf) + V[6] * (Pg))
    d.Qx[0] = Vx[0] * P11; d.Qy[0] = Vy[0] * P11; d.Qvx[0] = Ax[0] * P11; d.Qvy[0] = Ay[0] * P11;
    d.Qx[1] = SG_QCOL(Vx, 1, P12, P32, P42, P52, P62, P72);
    d.Qy[1] = SG_QCOL(Vy, 1, P12, P32, P42, P52, P62, P72);
    d.Qvx[1] = SG_QCOL(Ax, 1, P12, P32, P42, P52, P62, P72);
    d.Qvy[1] = SG_QCOL(Ay, 1, P12, P32, P42, P52, P62, P72);
    d.Qx[2] = SG_QCOL(Vx, 2, P13, P33, P43, P53, P63, P73);
    d.Qy[2] = SG_QCOL(Vy, 2, P13, P33, P43, P53, P63, P73);
    d.Qvx[2] = SG_QCOL(Ax, 2, P13, P33, P43, P53, P63, P73);
    d.Qvy[2] = SG_QCOL(Ay, 2, P13, P33, P43, P53, P63, P73);
    d.Qx[3] = SG_QCOL(Vx, 3, P14, P34, P44, P54, P64, P74);
    d.Qy[3] = SG_QCOL(Vy, 3, P14, P34, P44, P54, P64, P74);
    d.Qvx[3] = SG_QCOL(Ax, 3, P14, P34, P44, P54, P64, P74);
    d.Qvy[3] = SG_QCOL(Ay, 3, P14, P34, P44, P54, P64, P74);
#undef SG_QCOL
}

// value and d/ds of event `ev` at s in [0, 1] (dynamic_model.py:183-217 on the interpolant).  `ev` differs from lane
// to lane, so the planet / kind are picked with selects, not indexed accesses.
template <class TR>
__device__ __forceinline__ void event_g_dg(const DevCfg& c, const DenseQ<TR>& d, const double (&px)[TR::NP],
                                           const double (&py)[TR::NP], int ev, double s, double& g, double& dg) {
    const double h = d.h;
    const double x = fma(h * s, fma(s, fma(s, fma(s, d.Qx[3], d.Qx[2]), d.Qx[1]), d.Qx[0]), d.x0);
    const double y = fma(h * s, fma(s, fma(s, fma(s, d.Qy[3], d.Qy[2]), d.Qy[1]), d.Qy[0]), d.y0);
    const double xs = h * fma(s, fma(s, fma(s, 4.0 * d.Qx[3], 3.0 * d.Qx[2]), 2.0 * d.Qx[1]), d.Qx[0]);  // dx/ds
    const double ys = h * fma(s, fma(s, fma(s, 4.0 * d.Qy[3], 3.0 * d.Qy[2]), 2.0 * d.Qy[1]), d.Qy[0]);
    constexpr int NPL = TR::NEV - 3;  // planet-type events
    double pxs = px[0], pys = py[0];
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 1; p < TR::NP; ++p) {
            pxs = (ev == p) ? px[p] : pxs;
            pys = (ev == p) ? py[p] : pys;
        }
    }
    const double r = (TR::TASK == TASK_GOAL) ? c.r0 : (ev == 0 ? c.r0 : c.r1);
    const double ddx = pxs - x, ddy = pys - y;
    const double d2 = ddx * ddx + ddy * ddy;
    const double inv = fast_rsqrt(fmax(d2, 1e-300));
    const double dist = fast_sqrt0(d2);
    const double g_planet = dist - r, dg_planet = -(ddx * xs + ddy * ys) * inv;
    const double ax_ = c.half_ws - x, ay_ = c.half_ws - y;
    const double g_max = fmin(ax_, ay_), dg_max = (ax_ <= ay_) ? -xs : -ys;
    const double bx_ = c.half_ws + x, by_ = c.half_ws + y;
    const double g_min = fmin(bx_, by_), dg_min = (bx_ <= by_) ? xs : ys;
    const double w = fma(d.alpha * h, s, d.om0);
    const double g_om = c.om_max - fabs(w), dg_om = (w >= 0 ? -1.0 : 1.0) * (d.alpha * h);
    g = ev < NPL ? g_planet : (ev == NPL ? g_max : (ev == NPL + 1 ? g_min : g_om));
    dg = ev < NPL ? dg_planet : (ev == NPL ? dg_max : (ev == NPL + 1 ? dg_min : dg_om));
}

// Root of event `ev` inside the accepted step, as s in [0, 1].  What scipy.optimize.brentq returns for
// (t_old, t_new, xtol = rtol = 4 eps) is a point within 4 eps (1 + |t|) of a sign change; here a Newton iteration
// that never leaves the current sign-change bracket (bisection when it would) runs to the same tolerance.
// End-point conventions of brentq are kept: an exact zero at an end point is the root, equal signs are flagged.
template <class TR>
__device__ __forceinline__ double event_root(const DevCfg& c, const DenseQ<TR>& d, const double (&px)[TR::NP],
                                             const double (&py)[TR::NP], int ev, uint32_t& status) {
    double a = 0.0, b = 1.0, ga, gb, dga, dgb;
    event_g_dg<TR>(c, d, px, py, ev, 0.0, ga, dga);
    event_g_dg<TR>(c, d, px, py, ev, 1.0, gb, dgb);
    if (ga == 0) return 0.0;
    if (gb == 0) return 1.0;
    if (signbit(ga) == signbit(gb)) {
        status |= 2u;  // SG_ENV_BRENT_SIGN: the reference's brentq raises here
        return 1.0;
    }
    const double tol_s = 0.5 * (4 * kEps + 4 * kEps * (d.t_old + d.h)) * fast_rcp(d.h);
    double s = ga * fast_rcp(ga - gb);  // secant point of the bracket
    if (!(s > a && s < b)) s = 0.5;
#pragma unroll 1
    for (int it = 0; it < 60; ++it) {
        double g, dg;
        event_g_dg<TR>(c, d, px, py, ev, s, g, dg);
        if (g == 0) return s;
        if (signbit(g) == signbit(ga)) {
            a = s;
            ga = g;
        } else {
            b = s;
            gb = g;
        }
        double sn = s - g * fast_rcp(dg);
        if (!(sn > a && sn < b)) sn = 0.5 * (a + b);
        const double step = fabs(sn - s);
        s = sn;
        if (step <= tol_s || b - a <= tol_s) break;
    }
    return s;
}

// scipy:ivp.py:683-697: roots of all active events, the earliest one terminates; state := sol(t*)
template <class TR>
__device__ __forceinline__ void terminal_state(const DevCfg& c, const DenseQ<TR>& d, const double (&px)[TR::NP],
                                               const double (&py)[TR::NP], uint32_t active, double& x, double& y,
                                               double& th, double& vx, double& vy, double& om, uint32_t& status) {
    double best = INFINITY;
    // every lane walks its OWN active events (usually one)
#pragma unroll 1
    for (uint32_t mask = active; mask; mask &= mask - 1u) {
        const int ev = __ffs((int)mask) - 1;
        const double root = event_root<TR>(c, d, px, py, ev, status);
        best = fmin(best, root);
    }
    const double s = best, h = d.h, hs = h * s;
    x = fma(hs, fma(s, fma(s, fma(s, d.Qx[3], d.Qx[2]), d.Qx[1]), d.Qx[0]), d.x0);
    y = fma(hs, fma(s, fma(s, fma(s, d.Qy[3], d.Qy[2]), d.Qy[1]), d.Qy[0]), d.y0);
    vx = fma(hs, fma(s, fma(s, fma(s, d.Qvx[3], d.Qvx[2]), d.Qvx[1]), d.Qvx[0]), d.vx0);
    vy = fma(hs, fma(s, fma(s, fma(s, d.Qvy[3], d.Qvy[2]), d.Qvy[1]), d.Qvy[0]), d.vy0);
    om = fma(d.alpha, hs, d.om0);
    th = d.th0 + d.om0 * hs + 0.5 * d.alpha * hs * hs;
}

}  // namespace sg
