// sg_engine.cu — kernels and C ABI (include/spacegym_b200.h) of the B200-native batched
// Space-Gym stepping engine.  sm_100a only; no CPU path.
//
// Kernels
//   hot_kernel<TR, OT>     K = 1 (default step path): load, action translate, select_initial_step,
//                          ONE Runge-Kutta attempt, classify; envs that are done with the step get their
//                          observation / reward / TimeLimit epilogue here, the others go to device-side queues.
//   cont_kernel<TR, OT>    queue of envs that need more attempts (25 %): resumes solve_ivp's loop, epilogue.
//   tail_kernel<TR, OT>    queues of envs with a terminal event (dense output + Brent roots + epilogue) and of
//                          finished envs that need a new goal or an auto-reset; empties the queues.
//   step_kernel<TR, OT>    K fused env.step()s per launch, one thread per env: action translate ->
//                          RK45 with events -> observation -> reward (+ goal resample) -> TimeLimit
//                          -> auto-reset, state in registers across the K substeps; observation rows
//                          staged in shared memory and written with 128-bit coalesced stores.
//   reset_kernel<TR, OT>   (masked) env.reset(): hexagonal-tiling placement / Kepler draw from the
//                          per-env Philox stream.
//   vector_field_kernel    SpaceshipEnv.vector_field for every env.
//   dfma_peak_kernel       FP64 roof micro-benchmark.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/spacegym_b200.h"
#include "sg_device.cuh"
#include "sg_tail.cuh"

// SG_PART: the library can be built as six translation units compiled in parallel (spacegym_b200/build.py):
// part 0 = C ABI + dispatch + the 2-planet GoalEnv kernels, parts 1-5 = the kernels of one more env family each
// (Goal 3P, Goal 4P, Kepler, DoNotCrash, Goal 1P).  Without SG_PART the whole engine is one translation unit.
#if defined(SG_PART) && SG_PART != 0
#define SG_MAIN_PART 0
#else
#define SG_MAIN_PART 1
#endif

namespace sg {

#ifndef SG_KBLOCK
#define SG_KBLOCK 128
#endif
constexpr int kBlock = SG_KBLOCK;  // threads per CTA of every kernel but the tail kernel
// SG_CONT_GATHER: the hot kernel writes a 16-row continuation record (integrator state + action terms) and the
// continuation kernel gathers the scene (planets, goal / orbit) and the step's start position from the state block
// by env index, instead of a 26-row record with everything in coalesced scratch rows.
#ifndef SG_CONT_GATHER
#define SG_CONT_GATHER 1
#endif

// ---- state access ---------------------------------------------------------------------------
#ifndef SG_LD_MODE
#define SG_LD_MODE 0
#endif
// State rows are read once per kernel.  0 (default): plain loads.  1: ld.global.cs (streaming).  2:
// ld.global.L1::no_allocate.  Measured (8 batches of 262144 envs): 4.97e9 / 4.81e9 / 4.84e9 env-steps/s — the
// hot kernel alone is indifferent (39.4 / 40.4 / 39.5 us), the kernels that follow it gather by env index and
// like the rows cached.
__device__ __forceinline__ double ld_row(const double* p) {
#if SG_LD_MODE == 2
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#elif SG_LD_MODE == 1
    return __ldcs(p);
#else
    return *p;
#endif
}
template <class TR>
__device__ __forceinline__ void load_env(const DevCfg& c, long long i, EnvRegs<TR>& e) {
    const long long n = c.stride;
    e.x = ld_row(c.ship + i);
    e.y = ld_row(c.ship + n + i);
    e.th = ld_row(c.ship + 2 * n + i);
    e.vx = ld_row(c.ship + 3 * n + i);
    e.vy = ld_row(c.ship + 4 * n + i);
    e.om = ld_row(c.ship + 5 * n + i);
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) {
            e.px[p] = ld_row(c.planets + (2 * p) * n + i);
            e.py[p] = ld_row(c.planets + (2 * p + 1) * n + i);
        }
        e.gx = ld_row(c.goal + i);
        e.gy = ld_row(c.goal + n + i);
        e.ecc = e.ang = 0.0;
    } else {
        e.px[0] = 0.0;  // kepler.py:204-206: planet and border sit at the origin
        e.py[0] = 0.0;
        e.gx = e.gy = 0.0;
        e.ecc = (TR::TASK == TASK_KEPLER) ? ld_row(c.orbit + i) : 0.0;
        e.ang = (TR::TASK == TASK_KEPLER) ? ld_row(c.orbit + n + i) : 0.0;
    }
}

template <class TR>
__device__ __forceinline__ void store_ship(const DevCfg& c, long long i, const EnvRegs<TR>& e) {
    const long long n = c.stride;
    c.ship[i] = e.x;
    c.ship[n + i] = e.y;
    c.ship[2 * n + i] = e.th;
    c.ship[3 * n + i] = e.vx;
    c.ship[4 * n + i] = e.vy;
    c.ship[5 * n + i] = e.om;
}

template <class TR>
__device__ __forceinline__ void store_scene(const DevCfg& c, long long i, const EnvRegs<TR>& e, bool planets) {
    const long long n = c.stride;
    if (TR::TASK == TASK_GOAL) {
        if (planets) {
#pragma unroll
            for (int p = 0; p < TR::NP; ++p) {
                c.planets[(2 * p) * n + i] = e.px[p];
                c.planets[(2 * p + 1) * n + i] = e.py[p];
            }
        }
        c.goal[i] = e.gx;
        c.goal[n + i] = e.gy;
    } else if (planets) {
        c.orbit[i] = e.ecc;
        c.orbit[n + i] = e.ang;
    }
}

// copy the result of reset_env (sg_device.cuh) into the registers
template <class TR>
__device__ __forceinline__ void adopt_reset(const DevCfg& c, EnvRegs<TR>& e, const double* out) {
    e.x = out[0]; e.y = out[1]; e.th = out[2]; e.vx = out[3]; e.vy = out[4]; e.om = out[5];
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) {
            e.px[p] = out[6 + 2 * p];
            e.py[p] = out[7 + 2 * p];
        }
        e.gx = out[6 + 2 * TR::NP];
        e.gy = out[7 + 2 * TR::NP];
    } else if (TR::TASK == TASK_KEPLER && c.randomize) {
        e.ecc = out[6];
        e.ang = out[7];
    }
}

// Block-cooperative copy of the staged observation rows: 128-bit stores, fully coalesced.
template <typename OT, int D>
__device__ __forceinline__ void flush_obs(const OT* sobs, OT* obs, long long block_first, long long n) {
    long long rows = n - block_first;
    if (rows > kBlock) rows = kBlock;
    const long long total = rows * D;  // elements
    OT* dst = obs + block_first * D;   // 16-byte aligned: kBlock * D * sizeof(OT) is a multiple of 16
    constexpr int V = 16 / sizeof(OT);
    const long long nvec = total / V;
    const int4* s4 = reinterpret_cast<const int4*>(sobs);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (long long j = threadIdx.x; j < nvec; j += kBlock) d4[j] = s4[j];
    for (long long j = nvec * V + threadIdx.x; j < total; j += kBlock) dst[j] = sobs[j];
}

// Warp-cooperative copy of the warp's own 32 staged observation rows (contiguous in global memory):
// 128-bit coalesced stores without a block-wide barrier.
template <typename OT, int D>
__device__ __forceinline__ void flush_obs_warp(const OT* sobs, OT* obs, long long block_first, long long n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long first = block_first + 32 * warp;
    long long rows = n - first;
    if (rows <= 0) return;
    if (rows >= 32) {
        // full warp (every warp but possibly the last of the grid): compile-time trip count, 32-bit indices, all
        // shared-memory loads in flight before the first store
        constexpr int NVEC = 32 * D * (int)sizeof(OT) / 16;
        static_assert((32 * D * sizeof(OT)) % 16 == 0, "a warp's rows are a whole number of 16-byte vectors");
        constexpr int IT = (NVEC + 31) / 32;
        const int4* s4 = reinterpret_cast<const int4*>(sobs + 32 * warp * D);
        int4* d4 = reinterpret_cast<int4*>(obs + first * D);
        int4 v[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k)
            if (k * 32 + lane < NVEC) v[k] = s4[k * 32 + lane];
#pragma unroll
        for (int k = 0; k < IT; ++k)
            if (k * 32 + lane < NVEC) d4[k * 32 + lane] = v[k];
        return;
    }
    const long long total = rows * D;             // elements
    const OT* src = sobs + 32 * warp * D;          // 16-byte aligned: 32 * D * sizeof(OT) is a multiple of 16
    OT* dst = obs + first * D;
    constexpr int V = 16 / sizeof(OT);
    const long long nvec = total / V;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (long long j = lane; j < nvec; j += 32) d4[j] = s4[j];
    for (long long j = nvec * V + lane; j < total; j += 32) dst[j] = src[j];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- the hot kernel --------------------------------------------------------------------------
template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock)
step_kernel(const DevCfg c, const float* __restrict__ actions, int K, unsigned long long step0, OT* __restrict__ obs,
            double* __restrict__ reward, uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out) {
    constexpr int D = TR::D;
    __shared__ __align__(16) OT sobs[kBlock * D];
    __shared__ double sstat[4][kBlock / 32];
    const long long block_first = (long long)blockIdx.x * kBlock;
    const long long i = block_first + threadIdx.x;
    const bool valid = i < c.n;
    double st_ret = 0, st_len = 0, st_eps = 0, st_goal = 0;

    if (valid) {
        EnvRegs<TR> e;
        load_env<TR>(c, i, e);
        uint32_t elapsed = c.elapsed[i];
        double ep_ret = c.ep_return[i];
        uint32_t status = 0;
        uint32_t rng_ctr = 0;
        bool rng_open = false, scene_dirty = false, planets_dirty = false;
        Rng rng;
        const unsigned long long genv = (unsigned long long)(c.first_env + i);

        for (int k = 0; k < K; ++k) {
            float r0, r1;
            if (actions) {
                float2 a = reinterpret_cast<const float2*>(actions)[(long long)k * c.stride + i];
                r0 = a.x;
                r1 = a.y;
            } else {
                random_action(c.seed, genv, step0 + (unsigned long long)k, TR::DISC != 0, r0, r1);
            }
            const Act act = translate_action<TR>(c, r0, r1);
            const double lx = e.x, ly = e.y;  // env.last_xy (spaceship_env.py:74)
            double pd[TR::NP];
            bool dn = make_step<TR>(c, e, act, status, pd);
            const bool last = (k == K - 1);
            if (last && obs) make_obs<TR, OT>(c, e, sobs + threadIdx.x * D);  // before _reward: old goal (quirk 3)

            double r;
            bool reached = false;
            if (TR::TASK == TASK_GOAL) {
                double cur;
                r = goal_reward<TR>(c, e, lx, ly, pd, cur);
                reached = cur < c.goal_radius;
            } else if (TR::TASK == TASK_KEPLER) {
                r = kepler_reward<TR>(c, e, act);
            } else {
                r = c.const_reward;  // do_not_crash.py:47-48
            }
            bool trunc = false;
            elapsed += 1;
            if (c.max_episode_steps > 0 && elapsed >= (uint32_t)c.max_episode_steps) {  // gym TimeLimit
                trunc = !dn;
                dn = true;
            }
            if (reached || (dn && c.auto_reset)) {  // cold: needs the env's RNG stream
                if (!rng_open) {
                    rng_ctr = c.rng_ctr[i];
                    rng.open(c.seed, genv, rng_ctr);
                    rng_open = true;
                }
            }
            if (TR::TASK == TASK_GOAL && reached) {
                r += c.sparse;
                st_goal += 1;
            }
            // the new goal of an env that restarts in the same step would be overwritten by the reset: not drawn
            if (TR::TASK == TASK_GOAL && reached && !(dn && c.auto_reset)) {
                double gxy[2];
                rng.fresh_block();
                resample_goal<TR>(c, i, &rng, gxy, &status);
                e.gx = gxy[0];
                e.gy = gxy[1];
                scene_dirty = true;
            }
            ep_ret += r;
            if (dn) {
                st_ret += ep_ret;
                st_len += (double)elapsed;
                st_eps += 1;
                if (c.auto_reset) {  // gym.vector (0.21): the returned observation is the reset one
                    double out[8 + 2 * TR::NP];
                    rng.fresh_block();
                    reset_env<TR>(c, i, &rng, out, &status);
                    adopt_reset<TR>(c, e, out);
                    scene_dirty = planets_dirty = true;
                    elapsed = 0;
                    ep_ret = 0.0;
                    if (last && obs) make_obs<TR, OT>(c, e, sobs + threadIdx.x * D);
                }
            }
            if (reward) reward[(long long)k * c.stride + i] = r;
            if (done_out) done_out[(long long)k * c.stride + i] = (uint8_t)((dn ? 1 : 0) | ((c.pack_flags && trunc) ? 2 : 0));
            if (trunc_out) trunc_out[(long long)k * c.stride + i] = trunc ? 1 : 0;
        }
        store_ship<TR>(c, i, e);
        if (scene_dirty) store_scene<TR>(c, i, e, planets_dirty);
        c.elapsed[i] = elapsed;
        c.ep_return[i] = ep_ret;
        if (rng_open) c.rng_ctr[i] = rng.ctr;
        if (status) c.status[i] |= status;
    }
    if (obs) {
        __syncthreads();
        flush_obs<OT, D>(sobs, obs, block_first, c.n);
    }
    // episode statistics: warp shuffle -> shared -> one atomic per block and counter
    st_ret = warp_sum(st_ret);
    st_len = warp_sum(st_len);
    st_eps = warp_sum(st_eps);
    st_goal = warp_sum(st_goal);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sstat[0][warp] = st_ret;
        sstat[1][warp] = st_len;
        sstat[2][warp] = st_eps;
        sstat[3][warp] = st_goal;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) s += sstat[threadIdx.x][w];
        if (s != 0) atomicAdd(&c.stats[threadIdx.x], s);
    }
}

// =================================================================================================
// Split step path (K = 1): hot -> continuation -> tail, connected by global work queues.
//
// A 32-wide warp only runs at full width while its lanes do the same thing.  Per env step 79 % of the
// envs need exactly one Runge-Kutta step, 21 % need more, ~2 % hit a terminal event (dense output +
// Brent roots + reset) and ~0.3 % reach their goal (tile bookkeeping + RNG).  In one monolithic kernel
// nearly every warp pays for all of these paths.  Here the hot kernel runs the common path for every
// env (initial step selection + ONE attempt + epilogue) and appends the rest to compact queues, which
// the continuation and tail kernels process with densely packed warps.
// =================================================================================================
constexpr int kContFields = SG_CONT_GATHER ? 18 : 28;  // x y th vx vy om cth sth Ax Ay h_abs t (lx ly) rejected | fs alpha act_term | (scene[10])
constexpr int kEventFields = 25;  // x y th vx vy om h t_old Ax[7] Ay[7] lx ly active
constexpr uint32_t kPostReached = 1u << 30, kPostDone = 1u << 31, kPostIndex = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Queue atomics as inline PTX.  Written as atomicAdd() inside `if (lane == 0)`, the compiler substitutes its
// own warp-aggregation sequence (VOTEU / ATOMG / SHFL.IDX) whose shuffle consumes the result on the spot, so
// the L2 round trip is waited for where the atomic is issued (and a second atomic only starts after the
// first returned).  The asm form returns into a register that nothing reads until the slot is needed.
__device__ __forceinline__ uint32_t atom_add_u32(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long atom_add_u64(void* p, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}

// warp-aggregated queue append; must be called by all 32 lanes of the warp
__device__ __forceinline__ uint32_t queue_push(uint32_t* counter, bool need) {
    const unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atom_add_u32(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

// three appends at once: lane 0 issues the (up to) three atomics back to back, so their L2 round trips overlap
__device__ __forceinline__ void queue_push3(uint32_t* c0, bool n0, uint32_t* c1, bool n1, uint32_t* c2, bool n2,
                                            uint32_t& s0, uint32_t& s1, uint32_t& s2) {
    const unsigned m0 = __ballot_sync(0xffffffffu, n0), m1 = __ballot_sync(0xffffffffu, n1),
                   m2 = __ballot_sync(0xffffffffu, n2);
    const int lane = threadIdx.x & 31;
    uint32_t b0 = 0, b1 = 0, b2 = 0;
    if (lane == 0) {
        if (m0) b0 = atom_add_u32(c0, (uint32_t)__popc(m0));
        if (m1) b1 = atom_add_u32(c1, (uint32_t)__popc(m1));
        if (m2) b2 = atom_add_u32(c2, (uint32_t)__popc(m2));
    }
    const unsigned below = (1u << lane) - 1u;
    s0 = __shfl_sync(0xffffffffu, b0, 0) + (uint32_t)__popc(m0 & below);
    s1 = __shfl_sync(0xffffffffu, b1, 0) + (uint32_t)__popc(m1 & below);
    s2 = __shfl_sync(0xffffffffu, b2, 0) + (uint32_t)__popc(m2 & below);
}

// episode statistics: warp shuffle, then one atomic per warp and non-zero counter (no block barrier)
__device__ __forceinline__ void warp_stats_flush(const DevCfg& c, double (&st)[4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // most warps end no episode and reach no goal in a step: one vote instead of a 5-deep dependent
        // shuffle/add chain per counter (the chain is pure latency at the end of the warp's life)
        if (!__any_sync(0xffffffffu, st[k] != 0)) continue;
        const double v = warp_sum(st[k]);
        if (lane == k && v != 0) atomicAdd(&c.stats[k], v);
    }
}

template <int BLOCK>
__device__ __forceinline__ void block_stats_flush(const DevCfg& c, double (&st)[4]) {
    __shared__ double sstat[4][BLOCK / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double v = warp_sum(st[k]);
        if (lane == 0) sstat[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) s += sstat[threadIdx.x][w];
        if (s != 0) atomicAdd(&c.stats[threadIdx.x], s);
    }
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The kernels of a step (and of consecutive steps) form a chain on one stream.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, a kernel's CTAs may become resident while the previous kernel
// drains; pdl_wait() (griddepcontrol.wait) then blocks until the previous grid has completed and its writes are
// visible.  Everything a kernel reads that an earlier kernel of the chain wrote comes after pdl_wait().
// pdl_trigger() (griddepcontrol.launch_dependents) is issued once a CTA has no long work left.  Without the launch
// attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- -DSG_CHECK_QUEUES: queue invariants, counted on the device (see DevCfg) ----------------------------------------
#ifdef SG_CHECK_QUEUES
__device__ __forceinline__ void chk_fail(const DevCfg& c, int k) { atomicAdd(&c.chk_err[k], 1u); }
__device__ __forceinline__ void chk_visit(const DevCfg& c, long long i) { atomicAdd(&c.chk_visits[i], 1u); }
__device__ __forceinline__ void chk_slot(const DevCfg& c, bool pushed, uint32_t slot) {
    if (pushed && (long long)slot >= c.q_cap) chk_fail(c, 1);
}
__device__ __forceinline__ void chk_consume(const DevCfg& c, int q, long long i) {
    if (i < 0 || i >= c.n) {
        chk_fail(c, 4);
        return;
    }
    if (atomicExch(&c.chk_mark[(size_t)q * (size_t)c.stride + (size_t)i], c.chk_epoch) == c.chk_epoch) chk_fail(c, 3);
}
__global__ void chk_pre_kernel(const DevCfg c) {
    if (threadIdx.x < 11 && threadIdx.x != 4 && threadIdx.x != 5 && c.q_counts[threadIdx.x] != 0) chk_fail(c, 0);
}
__global__ void chk_post_kernel(const DevCfg c, uint32_t expected) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x)
        if (c.chk_visits[i] != expected) chk_fail(c, 2);
    if (blockIdx.x == 0 && threadIdx.x < 11 && threadIdx.x != 4 && threadIdx.x != 5 && c.q_counts[threadIdx.x] != 0)
        chk_fail(c, 5);
}
#define SG_CHK(stmt) stmt
#else
#define SG_CHK(stmt)
#endif

template <class TR>
__device__ __forceinline__ Act fetch_action(const DevCfg& c, const float* __restrict__ actions, long long i,
                                            unsigned long long step0) {
    float r0, r1;
    if (actions) {
        float2 a = reinterpret_cast<const float2*>(actions)[i];
        r0 = a.x;
        r1 = a.y;
    } else {
        random_action(c.seed, (unsigned long long)(c.first_env + i), step0, TR::DISC != 0, r0, r1);
    }
    return translate_action<TR>(c, r0, r1);
}

// Epilogue of an env whose physics is finished: observation (old goal, quirk 3), reward, TimeLimit,
// bookkeeping and outputs.  The goal resample / reset are NOT done here (post work).
template <class TR, typename OT>
__device__ __forceinline__ void finish_env(const DevCfg& c, long long i, EnvRegs<TR>& e, const Act& act, double lx,
                                           double ly, const double (&pd)[TR::NP], const double (&ipd)[TR::NP], bool dn,
                                           OT* obs_row, double cs, double sn, uint32_t elapsed0, double ep_ret0,
                                           double* __restrict__ reward, uint8_t* __restrict__ done_out,
                                           uint8_t* __restrict__ trunc_out, double (&st)[4], bool& reached, bool& done) {
    if (obs_row) make_obs_inv<TR, OT>(c, e, cs, sn, ipd, obs_row);
    double r;
    reached = false;
    if (TR::TASK == TASK_GOAL) {
        double cur;
        r = goal_reward<TR>(c, e, lx, ly, pd, cur);
        reached = cur < c.goal_radius;
        if (reached) {
            r += c.sparse;
            st[3] += 1;
        }
    } else if (TR::TASK == TASK_KEPLER) {
        r = kepler_reward<TR>(c, e, act);
    } else {
        r = c.const_reward;  // do_not_crash.py:47-48
    }
    uint32_t elapsed = elapsed0 + 1;
    bool trunc = false;
    if (c.max_episode_steps > 0 && elapsed >= (uint32_t)c.max_episode_steps) {  // gym TimeLimit
        trunc = !dn;
        dn = true;
    }
    const double ep_ret = ep_ret0 + r;
    if (dn) {
        st[0] += ep_ret;
        st[1] += (double)elapsed;
        st[2] += 1;
    }
    store_ship<TR>(c, i, e);
    c.elapsed[i] = elapsed;
    c.ep_return[i] = ep_ret;
    if (reward) reward[i] = r;
    if (done_out) done_out[i] = (uint8_t)((dn ? 1 : 0) | ((c.pack_flags && trunc) ? 2 : 0));
    if (trunc_out) trunc_out[i] = trunc ? 1 : 0;
    done = dn;
    SG_CHK(chk_visit(c, i);)
}

// Post work of a finished env: goal resample (goal.py:155-157) and/or auto-reset (gym.vector 0.21:
// the returned observation becomes the reset one), both from the env's Philox stream.
template <class TR, typename OT>
__device__ __forceinline__ void post_env(const DevCfg& c, long long i, bool reached, bool done, OT* __restrict__ obs) {
    Rng rng;
    rng.open(c.seed, (unsigned long long)(c.first_env + i), c.rng_ctr[i]);
    uint32_t status = 0;
    if (TR::TASK == TASK_GOAL && reached && !(done && c.auto_reset)) {
        double gxy[2];
        rng.fresh_block();
        resample_goal<TR>(c, i, &rng, gxy, &status);
        c.goal[i] = gxy[0];
        c.goal[c.stride + i] = gxy[1];
    }
    if (done && c.auto_reset) {
        EnvRegs<TR> e;
        load_env<TR>(c, i, e);
        double out[8 + 2 * TR::NP];
        rng.fresh_block();
        reset_env<TR>(c, i, &rng, out, &status);
        adopt_reset<TR>(c, e, out);
        store_ship<TR>(c, i, e);
        store_scene<TR>(c, i, e, true);
        c.elapsed[i] = 0;
        c.ep_return[i] = 0.0;
        if (obs) {
            OT row[TR::D];
            make_obs<TR, OT>(c, e, row);
#pragma unroll
            for (int j = 0; j < TR::D; ++j) obs[i * TR::D + j] = row[j];
        }
    }
    c.rng_ctr[i] = rng.ctr;
    if (status) c.status[i] |= status;
}

template <class TR>
__device__ __forceinline__ void save_cont(const DevCfg& c, uint32_t* q_idx, double* scr, uint32_t slot, long long i,
                                          const Integ<TR>& S, double lx, double ly, bool rejected,
                                          bool write_index = true) {
    const long long cap = c.q_cap;
    double* p = scr + slot;
    if (write_index) q_idx[slot] = (uint32_t)i;
    p[0 * cap] = S.x; p[1 * cap] = S.y; p[2 * cap] = S.th; p[3 * cap] = S.vx; p[4 * cap] = S.vy; p[5 * cap] = S.om;
    p[6 * cap] = S.cth; p[7 * cap] = S.sth; p[8 * cap] = S.Ax; p[9 * cap] = S.Ay; p[10 * cap] = S.h_abs;
    p[11 * cap] = S.t; p[14 * cap] = rejected ? 1.0 : 0.0;
    if (!SG_CONT_GATHER) {  // otherwise the continuation kernel re-reads the step's start position from the state block
        p[12 * cap] = lx; p[13 * cap] = ly;
    }
}

// the action terms (and, without SG_CONT_GATHER, the scene) next to the integrator state of a continuation record
template <class TR>
__device__ __forceinline__ void save_cont_scene(const DevCfg& c, double* scr, uint32_t slot, const EnvRegs<TR>& e,
                                                const Act& act) {
    const long long cap = c.q_cap;
    double* p = scr + slot;
    p[15 * cap] = act.fs; p[16 * cap] = act.alpha; p[17 * cap] = act.act_term;
    if (SG_CONT_GATHER) return;  // planets / goal / orbit: gathered from the state block by env index
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int q = 0; q < TR::NP; ++q) {
            p[(18 + 2 * q) * cap] = e.px[q];
            p[(19 + 2 * q) * cap] = e.py[q];
        }
        p[(18 + 2 * TR::NP) * cap] = e.gx;
        p[(19 + 2 * TR::NP) * cap] = e.gy;
    } else {
        p[18 * cap] = e.ecc;
        p[19 * cap] = e.ang;
    }
}

template <class TR>
__device__ __forceinline__ void load_cont_scene(const DevCfg& c, const double* scr, uint32_t j, long long i, EnvRegs<TR>& e,
                                                Act& act) {
    const long long cap = c.q_cap;
    const double* p = scr + j;
    act.fs = p[15 * cap]; act.alpha = p[16 * cap]; act.act_term = p[17 * cap];
    act.w = 0.0;
    e.x = e.y = e.th = e.vx = e.vy = e.om = 0.0;
    if (SG_CONT_GATHER) {  // the hot kernel does not touch the scene fields, and the ship rows only of finished envs
        const long long n = c.stride;
        if (TR::TASK == TASK_GOAL) {
#pragma unroll
            for (int q = 0; q < TR::NP; ++q) {
                e.px[q] = c.planets[(2 * q) * n + i];
                e.py[q] = c.planets[(2 * q + 1) * n + i];
            }
            e.gx = c.goal[i];
            e.gy = c.goal[n + i];
            e.ecc = e.ang = 0.0;
        } else {
            e.px[0] = e.py[0] = 0.0;
            e.gx = e.gy = 0.0;
            e.ecc = (TR::TASK == TASK_KEPLER) ? c.orbit[i] : 0.0;
            e.ang = (TR::TASK == TASK_KEPLER) ? c.orbit[n + i] : 0.0;
        }
        return;
    }
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int q = 0; q < TR::NP; ++q) {
            e.px[q] = p[(18 + 2 * q) * cap];
            e.py[q] = p[(19 + 2 * q) * cap];
        }
        e.gx = p[(18 + 2 * TR::NP) * cap];
        e.gy = p[(19 + 2 * TR::NP) * cap];
        e.ecc = e.ang = 0.0;
    } else {
        e.px[0] = e.py[0] = 0.0;
        e.gx = e.gy = 0.0;
        e.ecc = p[18 * cap];
        e.ang = p[19 * cap];
    }
}

template <class TR>
__device__ __forceinline__ void save_event(const DevCfg& c, uint32_t slot, long long i, const Integ<TR>& S, double h,
                                           const double (&Ax)[7], const double (&Ay)[7], uint32_t active, double lx,
                                           double ly) {
    const long long cap = c.q_cap;
    double* p = c.scr_event + slot;
    c.q_event[slot] = (uint32_t)i;
    p[0 * cap] = S.x; p[1 * cap] = S.y; p[2 * cap] = S.th; p[3 * cap] = S.vx; p[4 * cap] = S.vy; p[5 * cap] = S.om;
    p[6 * cap] = h; p[7 * cap] = S.t;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        p[(8 + k) * cap] = Ax[k];
        p[(15 + k) * cap] = Ay[k];
    }
    p[22 * cap] = lx; p[23 * cap] = ly; p[24 * cap] = (double)active;
}

enum { KIND_DONE = 0, KIND_CONT = 1, KIND_EVENT = 2, KIND_FINISH = 3, KIND_CONT_REJECTED = 4 };

#ifndef SG_HOT_MINB
#define SG_HOT_MINB 4
#endif
#ifndef SG_CONT_MINB
#define SG_CONT_MINB 3
#endif
// Grid sizes of the queue-driven kernels, in CTAs per SM (both walk their queue with a grid stride).  The queue
// lengths are only known on the device, so the grids are sized for the resident capacity, not for the worst
// case: a CTA that finds no work still has to wait for a slot with its registers (21 K / 15 K) next to the hot
// CTAs of the other env batches in flight.  Measured, 8 batches of 262144 envs: 8/8 -> 4.16e9, 6/2 -> 4.50e9,
// 3/2 -> 4.65e9, 3/1 -> 4.69e9, 2/1 -> 4.69e9 env-steps/s.
#ifndef SG_CONT_GRID
#define SG_CONT_GRID 3
#endif
// Continuation kernel, measured on one B200 (16 batches of 262144 envs / one handle, gpurun_out r2 dev runs):
// observation rows written by the warp together (SG_CONT_OBS_T): 6.05e9 vs 5.74e9 env-steps/s / 4.33e9 vs 4.34e9;
// the next chunk's ticket read at the bottom of the loop body (SG_CONT_DEFER_TICKET): neutral; 4 CTAs per SM at 128
// registers (184 B of spills) on top of both: 6.03e9 / 4.38e9.
#ifndef SG_CONT_DEFER_TICKET
#define SG_CONT_DEFER_TICKET 1
#endif
#ifndef SG_CONT_OBS_T
#define SG_CONT_OBS_T 1
#endif
#ifndef SG_TAIL_GRID
#define SG_TAIL_GRID 1
#endif
// One chunk of kBlock envs through the common path.  `publish` = true (fused persistent kernel): a
// continuation entry becomes visible to the in-kernel consumers by the release of its index word.
// ---- staged inputs of the persistent hot kernel ---------------------------------------------------------
// Every thread copies the fields of ITS OWN env of the next chunk into shared memory with cp.async while it
// integrates the current chunk, so the DRAM latency of the state load (19 % of the warp's life in the
// non-staged kernel) is off the critical path.  No block barrier is needed: a thread only reads what it copied.
template <class TR>
struct Stage {
    static constexpr int NF = (TR::TASK == TASK_GOAL ? 6 + 2 * TR::NP + 2 : 6 + 2) + 2;  // + ep_return + action
    double f[NF][kBlock];
    uint32_t elapsed[kBlock];
    uint32_t zero[kBlock];  // written once by its thread; see the queue atomic in hot_chunk
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class TR>
__device__ __forceinline__ void stage_prefetch(const DevCfg& c, const float* __restrict__ actions, long long i,
                                               Stage<TR>& s) {
    const int t = threadIdx.x;
    if (i < c.n) {
        const long long n = c.stride;
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) cp_async8(&s.f[k++][t], c.ship + r * n + i);
        if (TR::TASK == TASK_GOAL) {
#pragma unroll
            for (int r = 0; r < 2 * TR::NP; ++r) cp_async8(&s.f[k++][t], c.planets + r * n + i);
            cp_async8(&s.f[k++][t], c.goal + i);
            cp_async8(&s.f[k++][t], c.goal + n + i);
        } else {
            cp_async8(&s.f[k++][t], c.orbit + i);
            cp_async8(&s.f[k++][t], c.orbit + n + i);
        }
        cp_async8(&s.f[k++][t], c.ep_return + i);
        if (actions) cp_async8(&s.f[k++][t], reinterpret_cast<const float2*>(actions) + i);
        cp_async4(&s.elapsed[t], c.elapsed + i);
    }
}

template <class TR>
__device__ __forceinline__ void stage_load(const DevCfg& c, const Stage<TR>& s, bool have_actions, EnvRegs<TR>& e,
                                           float& r0, float& r1) {
    const int t = threadIdx.x;
    int k = 0;
    e.x = s.f[k++][t]; e.y = s.f[k++][t]; e.th = s.f[k++][t]; e.vx = s.f[k++][t]; e.vy = s.f[k++][t]; e.om = s.f[k++][t];
    if (TR::TASK == TASK_GOAL) {
#pragma unroll
        for (int p = 0; p < TR::NP; ++p) {
            e.px[p] = s.f[k++][t];
            e.py[p] = s.f[k++][t];
        }
        e.gx = s.f[k++][t];
        e.gy = s.f[k++][t];
        e.ecc = e.ang = 0.0;
    } else {
        e.px[0] = e.py[0] = 0.0;
        e.gx = e.gy = 0.0;
        e.ecc = s.f[k++][t];
        e.ang = s.f[k++][t];
    }
    k++;  // ep_return: read from the stage when the epilogue needs it
    r0 = r1 = 0.0f;
    if (have_actions) {
        const float2 a = *reinterpret_cast<const float2*>(&s.f[k][t]);
        r0 = a.x;
        r1 = a.y;
    }
}

template <class TR, typename OT, bool STAGED = false>
__device__ __forceinline__ void hot_chunk(const DevCfg& c, const float* __restrict__ actions, unsigned long long step0,
                                          OT* __restrict__ obs, double* __restrict__ reward,
                                          uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out,
                                          long long block_first, OT* sobs, double (&st)[4], bool publish,
                                          const Stage<TR>* stage = nullptr, double* stash = nullptr) {
    constexpr int D = TR::D;
    OT* const my_row = sobs + threadIdx.x * D;
    const long long i = block_first + threadIdx.x;
    const bool valid = i < c.n;
    int kind = KIND_DONE;
    bool need_post = false, reached = false, done = false, rejected = false;
    Integ<TR> S;
    double Ax[7], Ay[7], h = 0, lx = 0, ly = 0;
    double fx = 0, fy = 0, fth = 0, fvx = 0, fvy = 0, fom = 0, fcn = 1, fsn = 0, fpd[TR::NP], fipd[TR::NP];
    uint32_t active = 0;

    EnvRegs<TR> e;
    Act act;
    uint32_t elapsed0 = 0;
    double ep_ret0 = 0.0;
    if (STAGED) {
        float r0 = 0.0f, r1 = 0.0f;
        if (valid) stage_load<TR>(c, *stage, actions != nullptr, e, r0, r1);
        if (valid) {
            if (!actions) random_action(c.seed, (unsigned long long)(c.first_env + i), step0, TR::DISC != 0, r0, r1);
            act = translate_action<TR>(c, r0, r1);
        }
    } else if (valid) {
        // What only the epilogue needs (episode return / length so far, last_xy) is parked in the thread's own
        // shared-memory slots for the duration of the integration instead of in registers: the two counters go
        // global -> shared without touching a register (cp.async).  Held in registers they were spilled to local
        // memory and came back from L2 (a long-scoreboard wait at the reload).
        reinterpret_cast<volatile uint32_t*>(&stash[1 * kBlock + threadIdx.x])[1] = 0u;  // see the queue atomic below
        cp_async8(&stash[0 * kBlock + threadIdx.x], c.ep_return + i);
        cp_async4(&stash[1 * kBlock + threadIdx.x], c.elapsed + i);
        cp_async_commit();
        load_env<TR>(c, i, e);
        act = fetch_action<TR>(c, actions, i, step0);
        stash[2 * kBlock + threadIdx.x] = e.x;  // env.last_xy (spaceship_env.py:74)
        stash[3 * kBlock + threadIdx.x] = e.y;
    }
    if (valid) {
        double g[TR::NEV], pd[TR::NP];
        init_step<TR>(c, e, act, S, g, pd);
        // first _step_impl attempt (scipy:rk.py:111-143); t = 0
        double t_new = S.h_abs;
        if (t_new - c.T > 0) t_new = c.T;
        h = t_new;
        S.h_abs = fabs(h);
        double xn, yn, thn, vxn, vyn, omn, cn, sn, err2, ipd[TR::NP];
        rk_attempt<TR>(c, e, S, h, Ax, Ay, xn, yn, thn, vxn, vyn, omn, cn, sn, err2, pd, ipd);
        if (!controller(err2, rejected, t_new < c.T, S.h_abs)) {
            kind = KIND_CONT_REJECTED;  // the attempt loop is resumed from the same state with the smaller step
        } else {
            double gn[TR::NEV];
            event_values_from_pd<TR>(c, xn, yn, omn, pd, gn);
            active = active_events<TR>(g, gn);
            if (active) {
                kind = KIND_EVENT;
                // The 14 stage forces of an event lane wait for their queue slot in the lane's own (otherwise unused)
                // row of the observation staging area, not in registers across the epilogue of the finished lanes:
                // 48 bytes of spills -> none at 128 registers (+2 %).  Staging the continuation lanes' integrator
                // state the same way measured slower.  Needs a 14-double row: f64 observations of Goal 3P / 4P.
                if (sizeof(OT) == 8 && D >= 14) {
                    double* r = reinterpret_cast<double*>(sobs) + threadIdx.x * D;
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        r[k] = Ax[k];
                        r[7 + k] = Ay[k];
                    }
                }
            } else if (t_new < c.T) {
                kind = KIND_CONT;
                S.x = xn; S.y = yn; S.th = thn; S.vx = vxn; S.vy = vyn; S.om = omn;
                S.cth = cn; S.sth = sn; S.Ax = Ax[6]; S.Ay = Ay[6]; S.t = t_new;
            } else {
                kind = KIND_FINISH;
                fx = xn; fy = yn; fth = thn; fvx = vxn; fvy = vyn; fom = omn; fcn = cn; fsn = sn;
#pragma unroll
                for (int p = 0; p < TR::NP; ++p) {
                    fpd[p] = pd[p];
                    fipd[p] = ipd[p];
                }
            }
        }
    }
    if (!STAGED && !publish) pdl_trigger();  // the long part of this CTA is over: the next kernel of the chain may launch
    // queue appends at warp-converged points.  The continuation / event atomics are issued BEFORE the
    // epilogue of the finished lanes, whose arithmetic then covers the L2 round trip of the atomics.
    const bool is_cont = kind == KIND_CONT || kind == KIND_CONT_REJECTED;
    const unsigned m_cont = __ballot_sync(0xffffffffu, is_cont);
    const unsigned m_event = __ballot_sync(0xffffffffu, kind == KIND_EVENT);
    const int lane = threadIdx.x & 31;
    uint32_t opaque_zero = 0;
    if (STAGED) {  // the stage still holds this chunk's inputs: x, y are rows 0, 1, ep_return the row before the action
        opaque_zero = *reinterpret_cast<const volatile uint32_t*>(&stage->zero[threadIdx.x]);
        if (valid) {
            ep_ret0 = stage->f[Stage<TR>::NF - 2][threadIdx.x];
            elapsed0 = stage->elapsed[threadIdx.x];
            lx = stage->f[0][threadIdx.x];
            ly = stage->f[1][threadIdx.x];
        }
    } else {
        opaque_zero = reinterpret_cast<volatile uint32_t*>(&stash[1 * kBlock + threadIdx.x])[1];
        if (valid) {
            cp_async_wait<0>();
            ep_ret0 = stash[0 * kBlock + threadIdx.x];
            elapsed0 = *reinterpret_cast<const uint32_t*>(&stash[1 * kBlock + threadIdx.x]);
            lx = stash[2 * kBlock + threadIdx.x];
            ly = stash[3 * kBlock + threadIdx.x];
        }
    }
    // q_counts[0] (continuation) and [1] (event) are one aligned 64-bit word: one atomic reserves both.  The
    // operand carries a term that is zero at run time (a word this thread parked in shared memory) but
    // lane-dependent as far as ptxas can tell: with a provably warp-uniform operand it rewrites even an inline-PTX atom into its own
    // elect / ATOMG / SHFL.IDX sequence, and that shuffle waits for the L2 round trip right here instead of
    // after the epilogue below.
    unsigned long long b_both = 0;
    if (lane == 0 && (m_cont | m_event))
        b_both = atom_add_u64(c.q_counts, ((unsigned long long)__popc(m_cont) | ((unsigned long long)__popc(m_event) << 32)) +
                                              (unsigned long long)opaque_zero);
    if (kind == KIND_FINISH) {
        uint32_t status = 0;
        commit_state<TR>(e, fx, fy, fth, fvx, fvy, fom, status);
        finish_env<TR, OT>(c, i, e, act, lx, ly, fpd, fipd, false, obs ? my_row : nullptr, fcn, fsn,
                           elapsed0, ep_ret0, reward, done_out, trunc_out, st, reached, done);
        need_post = reached || (done && c.auto_reset);
        if (status) c.status[i] |= status;
        kind = KIND_DONE;
    }
    const unsigned below = (1u << lane) - 1u;
    b_both = __shfl_sync(0xffffffffu, b_both, 0);
    const uint32_t s_cont = (uint32_t)b_both + (uint32_t)__popc(m_cont & below);
    const uint32_t s_event = (uint32_t)(b_both >> 32) + (uint32_t)__popc(m_event & below);
    const uint32_t s_post = queue_push(&c.q_counts[2], need_post);
    SG_CHK(chk_slot(c, is_cont, s_cont); chk_slot(c, kind == KIND_EVENT, s_event); chk_slot(c, need_post, s_post);)
    if (is_cont) {
        save_cont<TR>(c, c.q_cont, c.scr_cont, s_cont, i, S, lx, ly, kind == KIND_CONT_REJECTED, !publish);
        save_cont_scene<TR>(c, c.scr_cont, s_cont, e, act);
        if (publish) st_release(c.q_cont + s_cont, (uint32_t)i);  // scratch row visible before "ready"
    }
    if (kind == KIND_EVENT) {
        if (sizeof(OT) == 8 && D >= 14) {
            const double* r = reinterpret_cast<const double*>(sobs) + threadIdx.x * D;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                Ax[k] = r[k];
                Ay[k] = r[7 + k];
            }
        }
        save_event<TR>(c, s_event, i, S, h, Ax, Ay, active, lx, ly);
    }
    if (need_post) c.q_post[s_post] = (uint32_t)i | (reached ? kPostReached : 0u) | (done ? kPostDone : 0u);
    if (obs) {  // rows of queued envs are rewritten by the later kernels
        __syncwarp();
        flush_obs_warp<OT, D>(sobs, obs, block_first, c.n);
    }
}

#ifdef SG_HOT_MAXREG  // experiment: register cap independent of the CTA count ptxas derives it from
#define SG_HOT_BOUNDS __maxnreg__(SG_HOT_MAXREG)
#else
#define SG_HOT_BOUNDS __launch_bounds__(kBlock, SG_HOT_MINB)
#endif
template <class TR, typename OT>
__global__ void SG_HOT_BOUNDS
hot_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
           uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out) {
    __shared__ __align__(16) OT sobs[kBlock * TR::D];
    __shared__ double stash[4 * kBlock];
    pdl_wait();  // the previous step's tail kernel has written the state and emptied the queues
    const unsigned long long step0 = *reinterpret_cast<const unsigned long long*>(c.q_counts + 4);
    double st[4] = {0, 0, 0, 0};
    hot_chunk<TR, OT>(c, actions, step0, obs, reward, done_out, trunc_out, (long long)blockIdx.x * kBlock, sobs, st, false,
                      nullptr, stash);
    warp_stats_flush(c, st);
}

template <class TR>
__device__ __forceinline__ void load_cont(const DevCfg& c, const double* scr, uint32_t j, long long i, const Act& act,
                                          Integ<TR>& S, double& lx, double& ly, bool& rejected) {
    const long long cap = c.q_cap;
    const double* p = scr + j;
    S.x = p[0 * cap]; S.y = p[1 * cap]; S.th = p[2 * cap]; S.vx = p[3 * cap]; S.vy = p[4 * cap];
    S.om = p[5 * cap]; S.cth = p[6 * cap]; S.sth = p[7 * cap]; S.Ax = p[8 * cap]; S.Ay = p[9 * cap];
    S.h_abs = p[10 * cap]; S.t = p[11 * cap];
    if (SG_CONT_GATHER) {  // env.last_xy = the position the step started from: still in the state block
        lx = c.ship[i]; ly = c.ship[c.stride + i];
    } else {
        lx = p[12 * cap]; ly = p[13 * cap];
    }
    rejected = p[14 * cap] != 0.0;
    S.alpha = (TR::STEER == STEER_ACC) ? act.alpha : 0.0;
    S.fs = act.fs;
}

// Pipelined form of the hot kernel: gridDim = resident CTAs, each walks the chunks blockIdx.x, + gridDim.x, ...
// Every thread copies the inputs of ITS env of the next chunk global -> shared (cp.async, two stage buffers)
// while it integrates the current one, so only the first chunk of a CTA waits for DRAM (in the one-chunk kernel
// that wait is 12 % of a warp's life).  A thread only ever reads shared memory it wrote itself: no block
// barrier anywhere.  Dynamic shared memory: 2 stages + the observation staging rows.
template <class TR, typename OT>
struct HotPipeSmem {
    Stage<TR> stages[2];
    alignas(16) OT sobs[kBlock * TR::D];
};
template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock, SG_HOT_MINB)
hot_staged_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
                  uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out, uint32_t n_chunks) {
    const unsigned long long step0 = *reinterpret_cast<const unsigned long long*>(c.q_counts + 4);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HotPipeSmem<TR, OT>& sm = *reinterpret_cast<HotPipeSmem<TR, OT>*>(smem_raw);
    sm.stages[0].zero[threadIdx.x] = 0u;
    sm.stages[1].zero[threadIdx.x] = 0u;
    uint32_t chunk = blockIdx.x;
    int cur = 0;
    if (chunk < n_chunks) stage_prefetch<TR>(c, actions, (long long)chunk * kBlock + threadIdx.x, sm.stages[0]);
    cp_async_commit();
    for (; chunk < n_chunks; chunk += gridDim.x, cur ^= 1) {
        const uint32_t next = chunk + gridDim.x;
        if (next < n_chunks) stage_prefetch<TR>(c, actions, (long long)next * kBlock + threadIdx.x, sm.stages[cur ^ 1]);
        cp_async_commit();
        cp_async_wait<1>();  // everything but the group just committed has landed: this chunk's inputs
        double st[4] = {0, 0, 0, 0};  // per chunk: nothing but the loop counters is live across chunks
        hot_chunk<TR, OT, true>(c, actions, step0, obs, reward, done_out, trunc_out, (long long)chunk * kBlock, sm.sobs, st,
                                false, &sm.stages[cur]);
        warp_stats_flush(c, st);
        __syncwarp();  // the warp's observation rows are flushed before its lanes stage the next chunk's rows
    }
    cp_async_wait<0>();
}

// One more Runge-Kutta attempt for continuation entry j of the first queue (valid = the entry exists).
// Same shape and register budget as the hot path.  Envs that need yet another attempt go to the second
// queue (generic loop kernel).
template <class TR, typename OT>
__device__ __forceinline__ void cont1_entry(const DevCfg& c, uint32_t j, bool valid, long long i, OT* __restrict__ obs,
                                            double* __restrict__ reward, uint8_t* __restrict__ done_out,
                                            uint8_t* __restrict__ trunc_out, double (&st)[4]) {
    int kind = KIND_DONE;
    bool need_post = false, reached = false, done = false, rejected = false;
    Integ<TR> S;
    EnvRegs<TR> e;
    Act act;
    double Ax[7], Ay[7], h = 0, lx = 0, ly = 0;
    uint32_t active = 0;
    if (valid) {
        const uint32_t elapsed0 = c.elapsed[i];
        const double ep_ret0 = c.ep_return[i];
        load_cont_scene<TR>(c, c.scr_cont, j, i, e, act);
        load_cont<TR>(c, c.scr_cont, j, i, act, S, lx, ly, rejected);
        double g[TR::NEV], pd[TR::NP];
        event_values(c, e, S.x, S.y, S.om, g, pd);
        // one pass of the attempt loop of _step_impl (scipy:rk.py:119-165)
        const double min_step = min_step_at(S.t);
        if (S.h_abs < min_step && !rejected) S.h_abs = min_step;
        if (S.h_abs < min_step) {
            kind = KIND_CONT;  // TOO_SMALL_STEP: let the generic kernel flag it
        } else {
            double t_new = S.t + S.h_abs;
            if (t_new - c.T > 0) t_new = c.T;
            h = t_new - S.t;
            S.h_abs = fabs(h);
            double xn, yn, thn, vxn, vyn, omn, cn, sn, err2, ipd[TR::NP];
            rk_attempt<TR>(c, e, S, h, Ax, Ay, xn, yn, thn, vxn, vyn, omn, cn, sn, err2, pd, ipd);
            if (!controller(err2, rejected, t_new < c.T, S.h_abs)) {
                kind = KIND_CONT;
            } else {
                rejected = false;
                double gn[TR::NEV];
                event_values_from_pd<TR>(c, xn, yn, omn, pd, gn);
                active = active_events<TR>(g, gn);
                if (active) {
                    kind = KIND_EVENT;
                } else if (t_new < c.T) {
                    kind = KIND_CONT;
                    S.x = xn; S.y = yn; S.th = thn; S.vx = vxn; S.vy = vyn; S.om = omn;
                    S.cth = cn; S.sth = sn; S.Ax = Ax[6]; S.Ay = Ay[6]; S.t = t_new;
                } else {
                    uint32_t status = 0;
                    commit_state<TR>(e, xn, yn, thn, vxn, vyn, omn, status);
                    OT row[TR::D];
                    finish_env<TR, OT>(c, i, e, act, lx, ly, pd, ipd, false, obs ? row : nullptr, cn, sn, elapsed0,
                                       ep_ret0, reward, done_out, trunc_out, st, reached, done);
                    if (obs) {
#pragma unroll
                        for (int k = 0; k < TR::D; ++k) obs[i * TR::D + k] = row[k];
                    }
                    need_post = reached || (done && c.auto_reset);
                    if (status) c.status[i] |= status;
                }
            }
        }
    }
    uint32_t s_cont, s_event, s_post;
    queue_push3(&c.q_counts[6], kind == KIND_CONT, &c.q_counts[1], kind == KIND_EVENT, &c.q_counts[2], need_post, s_cont,
                s_event, s_post);
    if (kind == KIND_CONT) {
        save_cont<TR>(c, c.q_cont2, c.scr_cont2, s_cont, i, S, lx, ly, rejected);
        save_cont_scene<TR>(c, c.scr_cont2, s_cont, e, act);
    }
    if (kind == KIND_EVENT) save_event<TR>(c, s_event, i, S, h, Ax, Ay, active, lx, ly);
    if (need_post) c.q_post[s_post] = (uint32_t)i | (reached ? kPostReached : 0u) | (done ? kPostDone : 0u);
}

template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock, SG_HOT_MINB)
cont1_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
             uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out) {
    const uint32_t count = c.q_counts[0];
    double st[4] = {0, 0, 0, 0};
    for (uint32_t base = blockIdx.x * kBlock; base < count; base += gridDim.x * kBlock) {
        const uint32_t j = base + threadIdx.x;
        const bool valid = j < count;
        const long long i = valid ? (long long)c.q_cont[j] : 0;
        cont1_entry<TR, OT>(c, j, valid, i, obs, reward, done_out, trunc_out, st);
    }
    block_stats_flush<kBlock>(c, st);
}

// Fused persistent form of hot_kernel + cont1_kernel (default).  gridDim = resident CTAs.  Phase 1: CTAs pull
// chunks of kBlock envs from a ticket counter and run the common path.  Phase 2: CTAs that run out of chunks
// pull batches of continuation entries — written by any CTA of this launch, published by the release of the
// entry's index word — so the second attempts fill the partial last wave of the grid instead of a separate,
// poorly occupied launch.  A CTA only ever waits for entries of chunks that a RUNNING CTA has claimed.
constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;
template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock, SG_HOT_MINB)
fused_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
             uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out, uint32_t n_chunks) {
    const unsigned long long step0 = *reinterpret_cast<const unsigned long long*>(c.q_counts + 4);
    __shared__ __align__(16) OT sobs[kBlock * TR::D];
    __shared__ uint32_t s_ticket;
    __shared__ double stash[4 * kBlock];
    double st[4] = {0, 0, 0, 0};
    for (;;) {  // phase 1
        if (threadIdx.x == 0) s_ticket = atomicAdd(&c.q_counts[7], 1u);
        __syncthreads();
        const uint32_t chunk = s_ticket;
        __syncthreads();
        if (chunk >= n_chunks) break;
        hot_chunk<TR, OT>(c, actions, step0, obs, reward, done_out, trunc_out, (long long)chunk * kBlock, sobs, st, true,
                          nullptr, stash);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&c.q_counts[8], 1u);  // chunks whose entries are all published
        }
    }
    uint32_t* q_idx = c.q_cont;
    const uint32_t* chunks_done = c.q_counts + 8;
    for (;;) {  // phase 2
        if (threadIdx.x == 0) s_ticket = atomicAdd(&c.q_counts[9], 1u);
        __syncthreads();
        const uint32_t batch = s_ticket;
        __syncthreads();
        const unsigned long long j64 = (unsigned long long)batch * kBlock + threadIdx.x;
        uint32_t i = kEmptySlot;
        if (j64 < (unsigned long long)c.q_cap) {
            for (;;) {
                i = ld_acquire(q_idx + j64);
                if (i != kEmptySlot) break;
                if (ld_acquire(chunks_done) >= n_chunks) {  // every producer has published: re-read once
                    i = ld_acquire(q_idx + j64);
                    break;
                }
                __nanosleep(64);
            }
        }
        const bool have = (i != kEmptySlot);
        if (!__syncthreads_or(have ? 1 : 0)) break;  // entries are dense: an empty batch is past the end
        if (have) q_idx[j64] = kEmptySlot;           // leave the queue empty for the next step
        cont1_entry<TR, OT>(c, (uint32_t)j64, have, (long long)i, obs, reward, done_out, trunc_out, st);
    }
    block_stats_flush<kBlock>(c, st);
}

// envs that still need more Runge-Kutta steps (rare): resume solve_ivp's loop until it finishes
template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock, SG_CONT_MINB)
cont_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
            uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out, int second_queue) {
    // The record of a step that hit a terminal event (29 doubles) waits for its queue slot in the thread's own
    // shared-memory slot, not in registers: the slot comes from a warp-convergent atomic after the loop body.
    __shared__ __align__(16) EventIO<TR> s_io[kBlock];
    static_assert(sizeof(EventIO<TR>) >= TR::D * sizeof(OT), "a record slot holds an observation row (SG_CONT_OBS_T)");
    __shared__ uint32_t s_zero[kBlock / 32];  // a run-time zero ptxas cannot see through (the ticket atomic below)
    if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile uint32_t*>(&s_zero[threadIdx.x >> 5]) = 0u;
    pdl_wait();
    const uint32_t count = second_queue ? c.q_counts[6] : c.q_counts[0];
    const uint32_t* q_in = second_queue ? c.q_cont2 : c.q_cont;
    const double* scr_in = second_queue ? c.scr_cont2 : c.scr_cont;
    double st[4] = {0, 0, 0, 0};
    // Work is handed out per WARP from a ticket counter: the entries of a warp take as long as its slowest lane (one
    // to four more Runge-Kutta steps), so a static split leaves the grid waiting for its unluckiest CTA.  The first
    // round needs no ticket (warp w of the grid takes chunk w).
    const uint32_t warps_in_grid = gridDim.x * (kBlock / 32);
    uint32_t chunk = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    for (; (unsigned long long)chunk * 32 < count;) {
        const uint32_t j = chunk * 32 + (threadIdx.x & 31);
        // the ticket of the warp's NEXT chunk: taken now (lane 0), first read at the bottom of the loop body, so the
        // atomic's L2 round trip is covered by the body instead of being waited for here
        // (a 64-bit atomic on the aligned pair q_counts[10..11], whose high word stays 0, with an operand that carries
        // a run-time zero from shared memory: ptxas rewrites an atomic add of a provably warp-uniform operand into
        // its own elect / ATOMG / SHFL.IDX sequence, and a 32-bit one with a lane-dependent operand into a five-step
        // shuffle scan; either shuffle waits for the round trip on the spot)
        uint32_t ticket = 0;
        if ((threadIdx.x & 31) == 0)
            ticket = (uint32_t)atom_add_u64(c.q_counts + 10, 1ull + *reinterpret_cast<volatile uint32_t*>(&s_zero[threadIdx.x >> 5]));
#if !SG_CONT_DEFER_TICKET
        chunk = warps_in_grid + __shfl_sync(0xffffffffu, ticket, 0);
#endif
        const bool valid = j < count;
        bool is_event = false, need_post = false, reached = false, done = false;
        long long i = 0;
        int row_env = -1;  // SG_CONT_OBS_T: env whose observation row is staged in this lane's slot
        Integ<TR> S;
        EventIO<TR>& io = s_io[threadIdx.x];
        double lx = 0, ly = 0;
        if (valid) {
            i = q_in[j];
            SG_CHK(chk_consume(c, 0, i);)
            EnvRegs<TR> e;
            Act act;
            load_cont_scene<TR>(c, scr_in, j, i, e, act);  // action terms from the record; planets / goal gathered from the state block
            bool rejected;
            const uint32_t elapsed0 = c.elapsed[i];
            const double ep_ret0 = c.ep_return[i];
            load_cont<TR>(c, scr_in, j, i, act, S, lx, ly, rejected);
            double g[TR::NEV], pd[TR::NP];
            event_values(c, e, S.x, S.y, S.om, g, pd);
            const int rc = rk_loop<TR>(c, e, S, g, rejected, &io, pd);
            if (rc == RK_EVENT) {
                is_event = true;
            } else {
                uint32_t status = (rc == RK_FAILED) ? 1u : 0u;
                commit_state<TR>(e, S.x, S.y, S.th, S.vx, S.vy, S.om, status);
                OT row[TR::D];
                double ipd[TR::NP];
#pragma unroll
                for (int q = 0; q < TR::NP; ++q) ipd[q] = fast_rcp(pd[q]);
                finish_env<TR, OT>(c, i, e, act, lx, ly, pd, ipd, false, obs ? row : nullptr, S.cth, S.sth, elapsed0,
                                   ep_ret0, reward, done_out, trunc_out, st, reached, done);
                if (obs) {
#if SG_CONT_OBS_T
                    // the row waits in the lane's (free: no event) record slot; the warp writes its rows together below
                    OT* srow = reinterpret_cast<OT*>(&io);
#pragma unroll
                    for (int k = 0; k < TR::D; ++k) srow[k] = row[k];
                    row_env = (int)i;
#else
#pragma unroll
                    for (int k = 0; k < TR::D; ++k) obs[i * TR::D + k] = row[k];
#endif
                }
                need_post = reached || (done && c.auto_reset);
                if (status) c.status[i] |= status;
            }
        }
        uint32_t s_unused, s_event, s_post;
        queue_push3(&c.q_counts[6], false, &c.q_counts[1], is_event, &c.q_counts[2], need_post, s_unused, s_event, s_post);
        SG_CHK(chk_slot(c, is_event, s_event); chk_slot(c, need_post, s_post);)
        if (is_event) {
            Integ<TR> So;
            So.x = io.x; So.y = io.y; So.th = io.th; So.vx = io.vx; So.vy = io.vy; So.om = io.om; So.t = io.t_old;
            save_event<TR>(c, s_event, i, So, io.h, io.Ax, io.Ay, io.active, lx, ly);
        }
        if (need_post) c.q_post[s_post] = (uint32_t)i | (reached ? kPostReached : 0u) | (done ? kPostDone : 0u);
#if SG_CONT_OBS_T
        // Observation rows of the warp's finished lanes, written by the whole warp: consecutive lanes write consecutive
        // elements of a row (a row is D contiguous values in global memory), 32 / D rows per store instruction, instead
        // of D stores per lane that each touch 32 different rows (one 32-byte sector per lane and store).
        if (obs) {
            __syncwarp();
            const EventIO<TR>* wio = s_io + (threadIdx.x & ~31);
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int q = 0; q < TR::D; ++q) {
                const int idx = q * 32 + lane, r = idx / TR::D, k = idx - r * TR::D;
                const int env_r = __shfl_sync(0xffffffffu, row_env, r);
                if (env_r >= 0) obs[(long long)env_r * TR::D + k] = reinterpret_cast<const OT*>(wio + r)[k];
            }
            __syncwarp();  // the slots are rewritten by the warp's next chunk
        }
#endif
#if SG_CONT_DEFER_TICKET
        chunk = warps_in_grid + __shfl_sync(0xffffffffu, ticket, 0);
#endif
    }
    pdl_trigger();
    warp_stats_flush(c, st);  // no block barrier: a warp leaves as soon as its own entries are done
}

// terminal events (dense output + Brent roots + epilogue + reset) and the post work of finished envs
#ifndef SG_TAIL_BLOCK
#define SG_TAIL_BLOCK 64
#endif
constexpr int kTailBlock = SG_TAIL_BLOCK;
template <class TR, typename OT>
__global__ void __launch_bounds__(kTailBlock)
tail_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
            uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out) {
    pdl_wait();
    const unsigned long long step0 = *reinterpret_cast<const unsigned long long*>(c.q_counts + 4);
    const uint32_t n_event = c.q_counts[1], n_post = c.q_counts[2];
    const uint32_t total = n_event + n_post;
    const long long cap = c.q_cap;
    double st[4] = {0, 0, 0, 0};
    // The event record (25 scratch rows) goes global -> shared with cp.async, every copy in flight at once, into
    // the thread's own EventIO slot; handle_events works on it through the pointer.  Filled through registers the
    // record lives on the local stack and the compiler pairs every load with its stack store, i.e. 25 serialised
    // memory round trips (30 % of this kernel's stall samples).
    __shared__ __align__(16) EventIO<TR> s_io[kTailBlock];
    __shared__ double s_tail[3][kTailBlock];  // lx, ly, active
    for (uint32_t base = blockIdx.x * kTailBlock; base < total; base += gridDim.x * kTailBlock) {
        const uint32_t j = base + threadIdx.x;
        long long post_i = 0;
        bool post_reached = false, post_done = false, do_post = false;
        if (j < n_event) {
            const double* p = c.scr_event + j;
            EventIO<TR>& io = s_io[threadIdx.x];
            cp_async8(&io.x, p + 0 * cap); cp_async8(&io.y, p + 1 * cap); cp_async8(&io.th, p + 2 * cap);
            cp_async8(&io.vx, p + 3 * cap); cp_async8(&io.vy, p + 4 * cap); cp_async8(&io.om, p + 5 * cap);
            cp_async8(&io.h, p + 6 * cap); cp_async8(&io.t_old, p + 7 * cap);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                cp_async8(&io.Ax[k], p + (8 + k) * cap);
                cp_async8(&io.Ay[k], p + (15 + k) * cap);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) cp_async8(&s_tail[k][threadIdx.x], p + (22 + k) * cap);
            cp_async_commit();
            const long long i = c.q_event[j];
            SG_CHK(chk_consume(c, 1, i);)
            EnvRegs<TR> e;
            load_env<TR>(c, i, e);
            const Act act = fetch_action<TR>(c, actions, i, step0);
            const uint32_t elapsed0 = c.elapsed[i];  // gathered with the state, used by the epilogue
            const double ep_ret0 = c.ep_return[i];
            cp_async_wait<0>();
            const double lx = s_tail[0][threadIdx.x], ly = s_tail[1][threadIdx.x];
            io.active = (uint32_t)s_tail[2][threadIdx.x];
            io.alpha = (TR::STEER == STEER_ACC) ? act.alpha : 0.0;
#pragma unroll
            for (int q = 0; q < TR::NP; ++q) {
                io.px[q] = e.px[q];
                io.py[q] = e.py[q];
            }
            handle_events<TR>(c, &io);
            uint32_t status = io.status;
            commit_state<TR>(e, io.x, io.y, io.th, io.vx, io.vy, io.om, status);
            double gd[TR::NEV], pd[TR::NP];
            event_values(c, e, e.x, e.y, e.om, gd, pd);
            double sn, cs;
            sincos_cold(e.th, &sn, &cs);
            OT row[TR::D];
            const bool want_row = obs && !c.auto_reset;  // with auto-reset the reset observation replaces it
            bool reached, done;
            double ipd[TR::NP];
#pragma unroll
            for (int q = 0; q < TR::NP; ++q) ipd[q] = fast_rcp(pd[q]);
            finish_env<TR, OT>(c, i, e, act, lx, ly, pd, ipd, true, want_row ? row : nullptr, cs, sn, elapsed0, ep_ret0,
                               reward, done_out, trunc_out, st, reached, done);
            if (want_row) {
#pragma unroll
                for (int k = 0; k < TR::D; ++k) obs[i * TR::D + k] = row[k];
            }
            if (status) c.status[i] |= status;
            post_i = i;
            post_reached = reached;
            post_done = done;
            do_post = true;
        } else if (j < total) {
            const uint32_t w = c.q_post[j - n_event];
            post_i = (long long)(w & kPostIndex);
            SG_CHK(chk_consume(c, 2, post_i);)
            post_reached = (w & kPostReached) != 0;
            post_done = (w & kPostDone) != 0;
            do_post = true;
        }
        // one call site for both kinds of entries: the reset code exists once (this kernel runs each
        // instruction once per warp, so its time is largely instruction-fetch latency)
        if (do_post) post_env<TR, OT>(c, post_i, post_reached, post_done, obs);
    }
    warp_stats_flush(c, st);
    // the last block to finish empties the queues for the next step
    __shared__ bool last;
    __syncthreads();  // every warp of the block is done before the block takes its ticket
    __threadfence();
    if (threadIdx.x == 0) last = (atomicAdd(&c.q_counts[3], 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {  // ... and advances the device-side step counter that keys the synthetic-action stream
        if (threadIdx.x < 4) c.q_counts[threadIdx.x] = 0;
        if (threadIdx.x == 5) c.q_counts[6] = 0;
        if (threadIdx.x >= 7 && threadIdx.x <= 10) c.q_counts[threadIdx.x] = 0;
        if (threadIdx.x == 4) *reinterpret_cast<unsigned long long*>(c.q_counts + 4) = step0 + 1;
    }
}


// =================================================================================================================
// Tail kernel, second form: latency first.  CTA = two warps working on the SAME 32 queue entries:
//   warp 0 ("E"): terminal events — dense output, roots, terminal state, reward / done / statistics;
//   warp 1 ("R"): whatever needs the env's random stream — the auto-reset of those envs (every event is terminal, so
//                 `done` is known before the root is), and for the post-queue entries the goal resample / reset.
// The two are independent: the reset state does not depend on where the episode ended, and an env that restarts in
// the same step draws no new goal (it would be overwritten).  The only hazard is that R overwrites scene fields E
// still has to read; E signals "my loads are done" on a named barrier (bar.arrive, non-blocking) and R waits on it
// before its stores.  The critical path of the kernel is max(E, R) instead of E + R.
// =================================================================================================================
constexpr int kTail2Block = 64;
#ifndef SG_TAIL2_MINB
#define SG_TAIL2_MINB 4
#endif
#ifndef SG_TAIL2_FENCE
#define SG_TAIL2_FENCE 0
#endif
template <class TR, typename OT>
__global__ void __launch_bounds__(kTail2Block, SG_TAIL2_MINB)
tail2_kernel(const DevCfg c, const float* __restrict__ actions, OT* __restrict__ obs, double* __restrict__ reward,
             uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out) {
    __shared__ double s_u[2 * kDrawBlocks * 32];  // R: the lanes' precomputed uniforms, element k of lane l at [k * 32 + l]
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    const unsigned long long step0 = *reinterpret_cast<const unsigned long long*>(c.q_counts + 4);
    const uint32_t n_event = c.q_counts[1], n_post = c.q_counts[2];
    const uint32_t total = n_event + n_post;
    const long long cap = c.q_cap;
    double st[4] = {0, 0, 0, 0};
    for (uint32_t base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
        const uint32_t j = base + lane;
        if (role == 0) {
            // ---------------- E: terminal events ----------------
            const bool mine = j < n_event;
            long long i = 0;
            EnvRegs<TR> e;
            DenseQ<TR> d;
            double Ax[7], Ay[7], lx = 0, ly = 0, ep_ret0 = 0;
            uint32_t active = 0, elapsed0 = 0;
            Act act;
            if (mine) {
                const double* p = c.scr_event + j;
                i = c.q_event[j];
                SG_CHK(chk_consume(c, 1, i);)
                d.x0 = p[0 * cap]; d.y0 = p[1 * cap]; d.th0 = p[2 * cap]; d.vx0 = p[3 * cap]; d.vy0 = p[4 * cap];
                d.om0 = p[5 * cap]; d.h = p[6 * cap]; d.t_old = p[7 * cap];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    Ax[k] = p[(8 + k) * cap];
                    Ay[k] = p[(15 + k) * cap];
                }
                lx = p[22 * cap]; ly = p[23 * cap];
                active = (uint32_t)p[24 * cap];
                load_env<TR>(c, i, e);
                act = fetch_action<TR>(c, actions, i, step0);
                elapsed0 = c.elapsed[i];
                ep_ret0 = c.ep_return[i];
            }
            bar_arrive(1, kTail2Block);  // every read of the old scene is done: R may overwrite it
            if (mine) {
                d.alpha = (TR::STEER == STEER_ACC) ? act.alpha : 0.0;
                dense_build<TR>(d, Ax, Ay);
                uint32_t status = 0;
                double x, y, th, vx, vy, om;
                terminal_state<TR>(c, d, e.px, e.py, active, x, y, th, vx, vy, om, status);
                commit_state<TR>(e, x, y, th, vx, vy, om, status);
                double gd[TR::NEV], pd[TR::NP], ipd[TR::NP];
                event_values_inv<TR>(c, e, e.x, e.y, e.om, gd, pd, ipd);
                // with auto-reset the reset observation replaces the terminal one in `obs`; the terminal row then goes
                // to the optional terminal-observation buffer (sg_set_terminal_obs)
                OT* const row_dst = c.auto_reset ? reinterpret_cast<OT*>(c.term_obs) : obs;
                const bool want_row = obs && row_dst;
                double sn = 0, cs = 1;
                if (want_row) sincos_cold(e.th, &sn, &cs);
                OT row[TR::D];
                // finish_env without the state stores an auto-reset would overwrite anyway
                if (want_row) make_obs_inv<TR, OT>(c, e, cs, sn, ipd, row);
                if (want_row && c.auto_reset) {
#pragma unroll
                    for (int k = 0; k < TR::D; ++k) row_dst[i * TR::D + k] = row[k];
                }
                double r;
                bool reached = false;
                if (TR::TASK == TASK_GOAL) {
                    double cur;
                    r = goal_reward<TR>(c, e, lx, ly, pd, cur);
                    reached = cur < c.goal_radius;
                    if (reached) {
                        r += c.sparse;
                        st[3] += 1;
                    }
                } else if (TR::TASK == TASK_KEPLER) {
                    r = kepler_reward<TR>(c, e, act);
                } else {
                    r = c.const_reward;
                }
                const uint32_t elapsed = elapsed0 + 1;
                const double ep_ret = ep_ret0 + r;
                st[0] += ep_ret;  // every event is terminal: the episode ends here (TimeLimit cannot truncate it)
                st[1] += (double)elapsed;
                st[2] += 1;
                if (reward) reward[i] = r;
                if (done_out) done_out[i] = 1;
                if (trunc_out) trunc_out[i] = 0;
                SG_CHK(chk_visit(c, i);)
                if (!c.auto_reset) {
                    store_ship<TR>(c, i, e);
                    c.elapsed[i] = elapsed;
                    c.ep_return[i] = ep_ret;
                    if (want_row) {
#pragma unroll
                        for (int k = 0; k < TR::D; ++k) obs[i * TR::D + k] = row[k];
                    }
                    if (TR::TASK == TASK_GOAL && reached) {  // rare configuration: the serial draw session
                        Rng rng;
                        rng.open(c.seed, (unsigned long long)(c.first_env + i), c.rng_ctr[i]);
                        double gxy[2];
                        resample_goal<TR>(c, i, &rng, gxy, &status);
                        c.goal[i] = gxy[0];
                        c.goal[c.stride + i] = gxy[1];
                        c.rng_ctr[i] = rng.ctr;
                    }
                }
                if (status) atomicOr(&c.status[i], status);
            }
        } else {
            // ---------------- R: resets and goal resamples ----------------
            long long i = 0;
            bool do_reset = false, do_goal = false;
            if (j < n_event) {
                i = c.q_event[j];
                do_reset = c.auto_reset != 0;
            } else if (j < total) {
                const uint32_t w = c.q_post[j - n_event];
                i = (long long)(w & kPostIndex);
                SG_CHK(chk_consume(c, 2, i);)
                const bool reached = (w & kPostReached) != 0, done = (w & kPostDone) != 0;
                do_reset = done && c.auto_reset;
                do_goal = TR::TASK == TASK_GOAL && reached && !do_reset;
            }
            const bool work = do_reset || do_goal;
            uint32_t ctr = 0;
            Tiles t;
            if (work) {
                ctr = c.rng_ctr[i];
                if (do_goal && TR::TILED) load_tiles(c, i, t);
            }
            if (__any_sync(0xffffffffu, work)) {
                if (work) philox_fill<draw_blocks<TR>()>(c.seed, (unsigned long long)(c.first_env + i), ctr, s_u + lane);
            }
            Draws rng;
            rng.open(s_u + lane);
            uint32_t status = 0;
            double out[8 + 2 * TR::NP];
            double gx = 0, gy = 0;
            if (do_reset) reset_env_t<TR>(c, rng, out, t, status);
            if (do_goal) {
                if (TR::TILED) find_new_goal_t<TR::NP>(c, t, rng, gx, gy, status);
                else sample_outside_one_planet(c, rng, c.planets[i], c.planets[c.stride + i], c.goal_radius, gx, gy, status);
            }
            bar_sync(1, kTail2Block);  // E has read what it needs of these envs
            if (do_reset) {
                EnvRegs<TR> e;
                e.ecc = e.ang = 0.0;
                if (TR::TASK == TASK_KEPLER && !c.randomize) {  // the reference orbit stays
                    e.ecc = c.orbit[i];
                    e.ang = c.orbit[c.stride + i];
                }
                e.px[0] = e.py[0] = 0.0;
                e.gx = e.gy = 0.0;
                adopt_reset<TR>(c, e, out);
                store_ship<TR>(c, i, e);
                if (TR::TASK == TASK_GOAL || (TR::TASK == TASK_KEPLER && c.randomize)) store_scene<TR>(c, i, e, true);
                if (TR::TILED) store_tiles(c, i, t, true);
                c.elapsed[i] = 0;
                c.ep_return[i] = 0.0;
                if (obs && c.term_obs && j >= n_event) {
                    // an env that hit its TimeLimit was finished by the hot / continuation kernel, which wrote its
                    // terminal observation into `obs`: keep that row before the reset observation replaces it
                    OT* t = reinterpret_cast<OT*>(c.term_obs);
#pragma unroll
                    for (int k = 0; k < TR::D; ++k) t[i * TR::D + k] = obs[i * TR::D + k];
                }
                if (obs) {
                    double sn, cs;
                    sincos_2pi(e.th, sn, cs);
                    OT row[TR::D];
                    make_obs_cs<TR, OT>(c, e, cs, sn, row);
#pragma unroll
                    for (int k = 0; k < TR::D; ++k) obs[i * TR::D + k] = row[k];
                }
            }
            if (do_goal) {
                c.goal[i] = gx;
                c.goal[c.stride + i] = gy;
                if (TR::TILED) store_tiles(c, i, t, false);
            }
            if (work) c.rng_ctr[i] = ctr + rng.blocks_used();
            if (status) atomicOr(&c.status[i], status);
        }
        __syncthreads();  // one barrier phase per sweep: E may not run ahead into the next sweep's arrive
    }
    pdl_trigger();
    warp_stats_flush(c, st);
    // the last block to finish empties the queues for the next step
    // (no __threadfence(): nothing the last block writes is read again inside this kernel — every block read the
    // counters and the step number before its sweep, and the next kernel starts behind the kernel boundary)
    __shared__ bool last;
    __syncthreads();
#if SG_TAIL2_FENCE
    __threadfence();
#endif
    if (threadIdx.x == 0) last = (atomicAdd(&c.q_counts[3], 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {  // ... and advances the device-side step counter that keys the synthetic-action stream
        if (threadIdx.x < 4) c.q_counts[threadIdx.x] = 0;
        if (threadIdx.x == 5) c.q_counts[6] = 0;
        if (threadIdx.x >= 7 && threadIdx.x <= 10) c.q_counts[threadIdx.x] = 0;
        if (threadIdx.x == 4) *reinterpret_cast<unsigned long long*>(c.q_counts + 4) = step0 + 1;
    }
}

#if SG_MAIN_PART
struct StepCounterPtrs {
    uint32_t* q[8];
};
__global__ void set_step_counter_kernel(StepCounterPtrs p, int n_sub, unsigned long long value) {
    if ((int)threadIdx.x < n_sub) *reinterpret_cast<unsigned long long*>(p.q[threadIdx.x] + 4) = value;
}
#endif

template <class TR, typename OT>
__global__ void __launch_bounds__(kBlock)
reset_kernel(const DevCfg c, const uint8_t* __restrict__ mask, OT* __restrict__ obs) {
    constexpr int D = TR::D;
    __shared__ double s_u[2 * kDrawBlocks * kBlock];  // per warp: 32 lanes x 32 draws, element k of lane l at [k * 32 + l]
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    EnvRegs<TR> e;
    load_env<TR>(c, i, e);
    if (!mask || mask[i]) {
        double* u = s_u + (threadIdx.x >> 5) * (2 * kDrawBlocks * 32) + (threadIdx.x & 31);
        const uint32_t ctr = c.rng_ctr[i];
        philox_fill<draw_blocks<TR>()>(c.seed, (unsigned long long)(c.first_env + i), ctr, u);
        Draws rng;
        rng.open(u);
        uint32_t status = 0;
        double out[8 + 2 * TR::NP];
        Tiles t;
        reset_env_t<TR>(c, rng, out, t, status);
        adopt_reset<TR>(c, e, out);
        store_ship<TR>(c, i, e);
        store_scene<TR>(c, i, e, true);
        if (TR::TILED) store_tiles(c, i, t, true);
        c.elapsed[i] = 0;
        c.ep_return[i] = 0.0;
        c.rng_ctr[i] = ctr + rng.blocks_used();
        if (status) c.status[i] |= status;
    }
    if (obs) {  // envs outside the mask report their current observation
        OT row[D];
        make_obs<TR, OT>(c, e, row);
#pragma unroll
        for (int j = 0; j < D; ++j) obs[i * D + j] = row[j];
    }
}

// SpaceshipEnv.vector_field (spaceship_env.py:96-100): [vx, vy, omega, ax, ay, alpha]
template <class TR>
__global__ void __launch_bounds__(kBlock)
vector_field_kernel(const DevCfg c, const float* __restrict__ actions, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    EnvRegs<TR> e;
    load_env<TR>(c, i, e);
    float2 a = reinterpret_cast<const float2*>(actions)[i];
    const Act act = translate_action<TR>(c, a.x, a.y);
    double sn, cs, ax, ay;
    sincos(e.th, &sn, &cs);
    accel<TR>(c, e, e.x, e.y, cs, sn, act.fs, ax, ay);
    double om = (TR::STEER == STEER_VEL) ? act.w : e.om;
    out[i * 6 + 0] = e.vx;
    out[i * 6 + 1] = e.vy;
    out[i * 6 + 2] = om;
    out[i * 6 + 3] = ax;
    out[i * 6 + 4] = ay;
    out[i * 6 + 5] = (TR::STEER == STEER_ACC) ? act.alpha : 0.0;
}

#if SG_MAIN_PART
__global__ void fill_orbit_kernel(double* orbit, long long n, double ecc, double ang) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        orbit[i] = ecc;
        orbit[n + i] = ang;
    }
}

// FP64 roof: 8 independent DFMA chains per thread, no memory traffic.
__global__ void dfma_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

#endif  // SG_MAIN_PART

}  // namespace sg

// =================================================================================================
// host side: C ABI
// =================================================================================================
using namespace sg;

struct sg_env {
    sg_config cfg;
    sg_layout lay;
    DevCfg dev;
    int device, sm_count;
    uint64_t seed;
    int64_t n, first_env;
    void* state;
    bool owns_state;
    double* d_stats;
    void* d_queues;  // work queues + scratch of the split step path
    uint32_t* d_chk;  // -DSG_CHECK_QUEUES: error counters [8] | visits [N] | marks [3][N]
    uint32_t chk_steps;
    size_t q_bytes_total;
    // split path: the envs are processed as `n_sub` independent sub-batches on internal streams, so that
    // the latency-bound continuation / tail kernels of one sub-batch overlap the hot kernel of another;
    // the whole fork / join is replayed as one CUDA graph per distinct set of argument pointers
    int n_sub;
    struct Sub {
        DevCfg dev;
        long long off, n;
        cudaStream_t stream;
        cudaEvent_t done;
    } sub[8];
    cudaEvent_t fork_ev;
    // the small latency-bound kernels of a step (continuation, tail) run on a high-priority stream so that
    // their blocks are scheduled ahead of the pending hot-kernel blocks of other handles' steps
    cudaStream_t hi_stream;
    cudaEvent_t ev_hot, ev_tail;
    int use_priority, use_fused, use_cont1, use_staged;
    void* term_obs = nullptr;  // sg_set_terminal_obs: [N][D] rows of the observation dtype of the step calls
    int use_pdl;      // SG_PDL (default 1): programmatic dependent launch along the kernel chain
    int tail_mode;    // SG_TAIL: 2 (default) = two-role latency-first tail kernel, 1 = the serial round-1 form
    int tail_grid;    // SG_TAIL_CTAS: CTAs per SM of the tail kernel's grid (32 queue entries per CTA and sweep)
    int timing;                 // SG_TIMING=1: CUDA events around every kernel of the split path
    cudaEvent_t tev[5];
    double t_acc[4];
    double q_acc[3];            // mean entries of the continuation / event / post queues
    long long t_n;
    struct GraphEntry {
        const void* key[6];
        cudaGraphExec_t exec;
        uint64_t stamp;
    } graphs[16];
    int n_graphs, graph_misses;
    uint64_t graph_clock;
    int use_graphs;
    uint64_t step_count;
    int64_t launches;
    std::string err;
    // staging for the *_host entry points
    cudaStream_t hstream;
    cudaEvent_t h_after;
    float* s_actions;
    size_t s_actions_bytes;
    void* s_block;         // device-side result block (obs | reward | flags)
    size_t s_block_bytes;
    double* s_reward;
    uint8_t* s_done;
    uint8_t* s_trunc;
    size_t s_k;
};

#if SG_MAIN_PART
static thread_local std::string g_create_error;

static int fail(sg_env* env, int code, const std::string& msg) {
    if (env) env->err = msg;
    else g_create_error = msg;
    return code;
}

#define SG_CUDA(env, call)                                                                         \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            return fail(env, SG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));  \
    } while (0)

extern "C" const char* sg_version(void) { return "spacegym_b200 0.1 (sm_100a)"; }

extern "C" const char* sg_last_error(const sg_env* env) { return env ? env->err.c_str() : g_create_error.c_str(); }

// hexagonal_tiling.py:161-174
static void tiling_rows_cols_a(int min_tiles, double world_size, int* r_out, int* c_out, double* a_out) {
    const double s3 = sqrt(3.0);
    double r_ = sqrt(72 * s3 * min_tiles - 6 * s3 + 12) / 12 - 1.0 / 4 + s3 / 12;
    int r = (int)ceil(r_), c;
    for (;;) {
        c = (int)floor(2 * s3 * r / 3 - 1.0 / 3 + s3 / 3);
        if (r * c >= min_tiles) break;
        ++r;
    }
    *r_out = r;
    *c_out = c;
    *a_out = 2 * s3 * world_size / (3 * (2 * r + 1));
}

extern "C" int sg_config_finalize(sg_config* cfg) {
    if (!cfg) return SG_ERR_INVALID;
    if (cfg->task == SG_TASK_GOAL) {
        // goal.py:10,62-72 fix these regardless of kwargs
        cfg->world_size = 3.0;
        cfg->step_size = 0.07;
        cfg->max_abs_vel_angle = 6.0;
        const int P = cfg->n_planets;
        if (P < 1 || P > SG_MAX_PLANETS) return SG_ERR_INVALID;
        if (P == 1) {
            // goal.py:79-82 (_init_one_planet): no tiling.  Upstream this configuration cannot take a step
            // (SURVEY §2.2-10); the semantics offered are the reference's with goal.py:134 and :153 repaired
            // (tests/golden/make_golden.py: make_goal_one_planet).
            cfg->rows = cfg->cols = 0;
            cfg->hex_a = cfg->hex_height = cfg->hex_width = cfg->tiling_width = 0;
            cfg->planets_radius = 0.8;
            cfg->goal_radius = 0.2;
            cfg->ship_radius = 0.2;
            return SG_OK;
        }
        const int n_objects = P + 2;
        const int min_tiles = (P == 2) ? n_objects : (int)ceil(n_objects / 0.6);  // hexagonal_tiling.py:26-29
        int r, c;
        double a;
        tiling_rows_cols_a(min_tiles, cfg->world_size, &r, &c, &a);
        if (r * c > 16 || c > SG_MAX_COLS) return SG_ERR_INVALID;
        cfg->rows = r;
        cfg->cols = c;
        cfg->hex_a = a;
        cfg->hex_height = a * sqrt(3.0);
        cfg->hex_width = 2 * a;
        cfg->tiling_width = 3 * a * (c - 1) / 2 + 2 * a;
        double pr = cfg->hex_height / 2;
        pr *= 0.75;  // PLANET_TILE_RATIO
        cfg->planets_radius = pr;
        cfg->goal_radius = cfg->ship_radius = pr / 2;
    } else if (cfg->task == SG_TASK_KEPLER) {
        cfg->n_planets = 2;
        cfg->world_size = 2 * cfg->kepler_border_radius;  // kepler.py:216
        cfg->max_abs_vel_angle = 6;
        cfg->rows = cfg->cols = 0;
        cfg->planets_radius = cfg->kepler_planet_radius;
        cfg->goal_radius = cfg->ship_radius = 0;
    } else if (cfg->task == SG_TASK_DONOTCRASH) {
        // do_not_crash.py:10-32: planet and border at the origin, world = the border's bounding square
        cfg->n_planets = 2;
        cfg->world_size = 2 * cfg->kepler_border_radius;
        cfg->max_abs_vel_angle = 5.0;
        cfg->step_size = 0.07;
        cfg->rows = cfg->cols = 0;
        cfg->planets_radius = cfg->kepler_planet_radius;
        cfg->goal_radius = cfg->ship_radius = 0;
    } else {
        return SG_ERR_INVALID;
    }
    if (cfg->ship_steering != SG_STEER_ACCELERATION && cfg->ship_steering != SG_STEER_VELOCITY) return SG_ERR_INVALID;
    return SG_OK;
}

extern "C" int sg_config_from_id(const char* id, sg_config* cfg) {
    if (!id || !cfg) return SG_ERR_INVALID;
    memset(cfg, 0, sizeof(*cfg));
    cfg->ship_mass = 1;
    cfg->max_thruster_force = 0.05;
    cfg->max_episode_steps = 500;
    cfg->auto_reset = 1;
    cfg->total_planets_mass = 1e9;  // goal.py:14
    cfg->kepler_planet_mass = 6e8;  // kepler.py:204
    cfg->kepler_planet_radius = 0.2;
    cfg->kepler_border_radius = 3.0;
    cfg->ship_moi = 0.01;
    cfg->max_engine_force = 0.4;
    cfg->ship_steering = SG_STEER_VELOCITY;
    int P = 0, disc = 0;
    if (!strcmp(id, "GoalContinuous2P-v0")) P = 2;
    else if (!strcmp(id, "GoalContinuous3P-v0")) P = 3;
    else if (!strcmp(id, "GoalContinuous4P-v0")) P = 4;
    else if (!strcmp(id, "GoalDiscrete2-v0")) P = 2, disc = 1;  // keyboard_agent.py:29-74
    else if (!strcmp(id, "GoalDiscrete3-v0")) P = 3, disc = 1;
    else if (!strcmp(id, "GoalDiscrete4-v0")) P = 4, disc = 1;
    if (P) {  // gym_space/__init__.py:26-72
        cfg->task = SG_TASK_GOAL;
        cfg->n_planets = P;
        cfg->discrete = disc;
        cfg->survival_reward_scale = 0.2;
        cfg->goal_vel_reward_scale = 5.0;
        cfg->safety_reward_scale = 10.0;
        cfg->goal_sparse_reward = 5.0;
        cfg->danger_zone = 0.25;
        if (disc) cfg->max_engine_force = 1;
        return sg_config_finalize(cfg);
    }
    if (!strcmp(id, "DoNotCrashContinuous-v0") || !strcmp(id, "DoNotCrashDiscrete-v0")) {
        // gym_space/__init__.py:5-15, do_not_crash.py:10-32.  Upstream omits ShipParams.steering (TypeError);
        // default here: acceleration steering (the class sets moi = 0.05, which only that mode uses).
        cfg->task = SG_TASK_DONOTCRASH;
        cfg->discrete = !strcmp(id, "DoNotCrashDiscrete-v0");
        cfg->max_episode_steps = 300;
        cfg->ship_steering = SG_STEER_ACCELERATION;
        cfg->ship_moi = 0.05;
        cfg->max_engine_force = 0.3;
        cfg->kepler_planet_radius = 0.25;
        cfg->kepler_border_radius = 1.0;
        return sg_config_finalize(cfg);
    }
    cfg->task = SG_TASK_KEPLER;  // gym_space/__init__.py:76-146
    cfg->rad_penalty_C = 2;
    cfg->numerator_C = 0.01;
    cfg->act_penalty_C = 0.5;
    cfg->step_size = 0.07;
    cfg->ref_orbit_a = 1.2;
    if (!strcmp(id, "KeplerCircleOrbit-v0")) {
        cfg->ref_orbit_eccentricity = 0;
        cfg->ref_orbit_angle = 0;
    } else if (!strcmp(id, "KeplerDiscrete-v0")) {  // keyboard_agent.py:10-27
        cfg->discrete = 1;
        cfg->max_episode_steps = 0;  // registered there without a TimeLimit
    } else if (!strcmp(id, "KeplerEllipseEasy-v0")) {
        cfg->ref_orbit_eccentricity = 0.5;
        cfg->ref_orbit_angle = 0.8;
    } else if (!strcmp(id, "KeplerEllipseHard-v0")) {
        cfg->ref_orbit_eccentricity = 0.725;
        cfg->ref_orbit_angle = 3.925;
    } else if (!strcmp(id, "KeplerRandomOrbits-v0")) {
        cfg->randomize = 1;
        cfg->ref_orbit_eccentricity = 0.5;  // kepler.py:193-194 until the first reset
        cfg->ref_orbit_angle = 3.75;
    } else {
        return SG_ERR_UNKNOWN_ID;
    }
    return sg_config_finalize(cfg);
}

extern "C" int sg_obs_dim(const sg_config* cfg) {
    if (!cfg) return SG_ERR_INVALID;
    if (cfg->task == SG_TASK_GOAL) return 9 + 2 * cfg->n_planets;
    return cfg->task == SG_TASK_KEPLER ? 10 : 7;  // kepler.py:172-187 appends the orbit; do_not_crash.py has no lidar / goal
}

extern "C" int sg_state_layout(const sg_config* cfg, int64_t n, sg_layout* out) {
    if (!cfg || !out || n <= 0) return SG_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->n_envs = n;
    int64_t off = 0;
    auto take = [&](int64_t rows, int64_t elem) {
        int64_t at = off;
        off += ((rows * n * elem + 255) / 256) * 256;
        return at;
    };
    const bool goal = cfg->task == SG_TASK_GOAL;
    out->ship = take(6, 8);
    out->planets = take(goal ? 2 * cfg->n_planets : 0, 8);
    out->goal = take(goal ? 2 : 0, 8);
    out->orbit = take(goal ? 0 : 2, 8);
    out->col_shift = take(goal ? cfg->cols : 0, 8);
    out->ep_return = take(1, 8);
    out->free_tiles = take(goal ? SG_FREE_WORDS : 0, 8);
    out->tile_word = take(goal ? 1 : 0, 4);
    out->elapsed = take(1, 4);
    out->rng_ctr = take(1, 4);
    out->status = take(1, 4);
    out->total_bytes = off;
    return SG_OK;
}

static size_t queue_bytes(long long n) {
    size_t b = 256 + 4 * ((((size_t)n * 4 + 255) / 256) * 256) +
               (size_t)(2 * kContFields + kEventFields) * (size_t)n * sizeof(double);
    return ((b + 255) / 256) * 256;
}

static void build_dev(sg_env* h) {
    const sg_config& c = h->cfg;
    DevCfg& d = h->dev;
    memset(&d, 0, sizeof(d));
    char* base = (char*)h->state;
    d.ship = (double*)(base + h->lay.ship);
    d.planets = (double*)(base + h->lay.planets);
    d.goal = (double*)(base + h->lay.goal);
    d.orbit = (double*)(base + h->lay.orbit);
    d.col_shift = (double*)(base + h->lay.col_shift);
    d.ep_return = (double*)(base + h->lay.ep_return);
    d.free_tiles = (unsigned long long*)(base + h->lay.free_tiles);
    d.tile_word = (uint32_t*)(base + h->lay.tile_word);
    d.elapsed = (uint32_t*)(base + h->lay.elapsed);
    d.rng_ctr = (uint32_t*)(base + h->lay.rng_ctr);
    d.status = (uint32_t*)(base + h->lay.status);
    d.stats = h->d_stats;
    d.stride = h->n;
    d.n = h->n;
    d.first_env = h->first_env;
    d.seed = h->seed;
    d.max_episode_steps = c.max_episode_steps;
    d.randomize = c.randomize;
    d.auto_reset = c.auto_reset;
    d.rows = c.rows;
    d.cols = c.cols;
    d.T = c.step_size;
    d.half_ws = c.world_size / 2;
    d.two_over_ws = 2 / c.world_size;
    d.om_max = c.max_abs_vel_angle;
    d.inv_mass = 1.0 / c.ship_mass;
    d.moi = c.ship_moi;
    d.max_engine = c.max_engine_force;
    d.max_thruster = c.max_thruster_force;
    d.max_engine_f32 = (float)c.max_engine_force;
    d.max_thruster_f32 = (float)c.max_thruster_force;
    if (c.task == SG_TASK_GOAL) {
        d.gm0 = kG * c.ship_mass * (c.total_planets_mass / c.n_planets);  // helpers.py:33, goal.py:43
        d.r0 = c.planets_radius;
        d.r1 = 0;
    } else {
        d.gm0 = kG * c.ship_mass * c.kepler_planet_mass;
        d.r0 = c.kepler_planet_radius;
        d.r1 = c.kepler_border_radius;
    }
    d.gmm = d.gm0 * d.inv_mass;
    d.survival = c.survival_reward_scale;
    d.goal_vel = c.goal_vel_reward_scale;
    d.safety = c.safety_reward_scale;
    d.sparse = c.goal_sparse_reward;
    d.danger_zone = c.danger_zone;
    d.goal_radius = c.goal_radius;
    d.ship_radius = c.ship_radius;
    d.ws = c.world_size;
    d.hex_a = c.hex_a;
    d.hex_height = c.hex_height;
    d.hex_width = c.hex_width;
    d.free_x = c.world_size - c.tiling_width;
    d.orbit_a = c.ref_orbit_a;
    d.numC = c.numerator_C;
    d.radC = c.rad_penalty_C;
    d.actC = c.act_penalty_C;
    d.actC_f32 = (float)c.act_penalty_C;
    if (c.task == SG_TASK_DONOTCRASH) {
        d.kep_lo = c.kepler_planet_radius + 0.2;  // do_not_crash.py:36-38
        d.kep_hi = c.kepler_border_radius - 0.15;
        d.reset_v_std = 0.07;                     // do_not_crash.py:41-43
        d.reset_om_div = 3;
    } else {
        d.kep_lo = c.kepler_planet_radius + 0.5;  // kepler.py:235-237
        d.kep_hi = c.kepler_border_radius - 0.5;
        d.reset_v_std = 0.05;                     // kepler.py:262-265
        d.reset_om_div = 5;
    }
    d.const_reward = 100.0 / 300;  // do_not_crash.py:11,47-48 (the class attribute, not the TimeLimit in force)
    // sub-batches: same constants, base pointers shifted by the sub-batch offset, own queue block
    char* q = (char*)h->d_queues;
    for (int s = 0; s < h->n_sub; ++s) {
        sg_env::Sub& sb = h->sub[s];
        DevCfg& x = sb.dev;
        x = d;
        x.n = sb.n;
        x.first_env = h->first_env + sb.off;
        x.ship += sb.off; x.planets += sb.off; x.goal += sb.off; x.orbit += sb.off; x.col_shift += sb.off;
        x.ep_return += sb.off; x.free_tiles += sb.off; x.tile_word += sb.off; x.elapsed += sb.off;
        x.rng_ctr += sb.off; x.status += sb.off;
        // queue block: counts (256 B) | cont idx | event idx | post idx | cont scratch | event scratch
        const size_t n = (size_t)sb.n, idx_bytes = ((n * 4 + 255) / 256) * 256;
        x.q_counts = (uint32_t*)q;
        x.q_cont = (uint32_t*)(q + 256);
        x.q_event = (uint32_t*)(q + 256 + idx_bytes);
        x.q_post = (uint32_t*)(q + 256 + 2 * idx_bytes);
        x.q_cont2 = (uint32_t*)(q + 256 + 3 * idx_bytes);
        x.scr_cont = (double*)(q + 256 + 4 * idx_bytes);
        x.scr_event = x.scr_cont + (size_t)kContFields * n;
        x.scr_cont2 = x.scr_event + (size_t)kEventFields * n;
        x.q_cap = sb.n;
#ifdef SG_CHECK_QUEUES
        x.chk_err = h->d_chk;
        x.chk_visits = h->d_chk ? h->d_chk + 8 + sb.off : nullptr;
        x.chk_mark = h->d_chk ? h->d_chk + 8 + h->n + sb.off : nullptr;  // [3][stride], stride = envs of the handle
        x.chk_epoch = 0;
#endif
        q += queue_bytes(sb.n);
    }
    h->dev.q_counts = h->sub[0].dev.q_counts;
}

extern "C" int sg_create(const sg_config* cfg, int64_t n_envs, int device, uint64_t seed, int64_t first_env_index,
                         void* d_state, sg_env** out) {
    if (!cfg || !out || n_envs <= 0) return fail(nullptr, SG_ERR_INVALID, "sg_create: bad arguments");
    if (n_envs >= (1ll << 30)) return fail(nullptr, SG_ERR_INVALID, "sg_create: at most 2^30 - 1 envs per handle");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, SG_ERR_NO_DEVICE, "sg_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, SG_ERR_INVALID, "sg_create: bad device index");
    sg_config c = *cfg;
    if (sg_config_finalize(&c) != SG_OK) return fail(nullptr, SG_ERR_INVALID, "sg_create: unsupported configuration");
    SG_CUDA(nullptr, cudaSetDevice(device));
    sg_env* h = new sg_env();
    h->sm_count = 148;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    h->cfg = c;
    h->device = device;
    h->seed = seed;
    h->n = n_envs;
    h->first_env = first_env_index;
    h->step_count = 0;
    h->launches = 0;
    h->d_queues = nullptr;
    h->d_chk = nullptr;
    h->chk_steps = 0;
    h->d_stats = nullptr;
    h->n_sub = 1;
    h->n_graphs = 0;
    h->graph_misses = 0;
    h->graph_clock = 0;
    h->fork_ev = nullptr;
    h->hi_stream = nullptr;
    h->ev_hot = h->ev_tail = nullptr;
    h->use_priority = 0;
    memset(h->sub, 0, sizeof(h->sub));
    memset(h->graphs, 0, sizeof(h->graphs));
    h->hstream = nullptr;
    h->s_actions = nullptr;
    h->s_actions_bytes = 0;
    h->s_block = nullptr;
    h->s_block_bytes = 0;
    h->h_after = nullptr;
    h->s_reward = nullptr;
    h->s_done = h->s_trunc = nullptr;
    h->s_k = 0;
    sg_state_layout(&c, n_envs, &h->lay);
    h->owns_state = (d_state == nullptr);
    h->state = d_state;
    cudaError_t err = cudaSuccess;
    if (h->owns_state) err = cudaMalloc(&h->state, (size_t)h->lay.total_bytes);
    if (err == cudaSuccess) err = cudaMalloc((void**)&h->d_stats, 4 * sizeof(double));
    {   // sub-batches of the split path: multiples of the block size, at least 32768 envs each
        // Measured on B200 (profiles/r1_notes.md, tools/bench_subbatches.py), one handle stepped back to back:
        // 262144 envs 3.34 / 3.48 / 3.44 / 3.41 G env-steps/s for 1 / 2 / 3 / 4 concurrent sub-batches (each still
        // walks the same three-kernel latency chain), 1048576 envs 4.58 / 5.29 / 5.90 / 5.90 G/s (the long thin
        // tail kernel of one sub-batch hides under the hot kernel of another).  Default: one below 3 x 262144
        // envs, then one sub-batch per ~350 k envs up to four; SG_SUBBATCHES overrides.
        int want = n_envs < 786432 ? 1 : (n_envs < 1048576 ? 2 : (n_envs < 2097152 ? 3 : 4));
        const char* env_sub = getenv("SG_SUBBATCHES");
        if (env_sub && atoi(env_sub) > 0) want = atoi(env_sub) > 8 ? 8 : atoi(env_sub);
        const char* env_timing = getenv("SG_TIMING");
        if (env_timing && atoi(env_timing)) want = 1;  // the per-kernel timing aid synchronises inside the step
        h->n_sub = want;
        const char* env_graph = getenv("SG_GRAPHS");
        h->use_graphs = env_graph ? atoi(env_graph) : 1;
        long long per = ((n_envs / want + kBlock - 1) / kBlock) * kBlock, off = 0;
        size_t q_bytes = 0;
        for (int s = 0; s < want; ++s) {
            long long cnt = (s == want - 1) ? n_envs - off : per;
            if (cnt <= 0) { h->n_sub = s; break; }
            h->sub[s].off = off;
            h->sub[s].n = cnt;
            h->sub[s].stream = nullptr;
            h->sub[s].done = nullptr;
            off += cnt;
            q_bytes += queue_bytes(cnt);
        }
        if (err == cudaSuccess) err = cudaMalloc(&h->d_queues, q_bytes);
        if (err == cudaSuccess) err = cudaMemset(h->d_queues, 0, q_bytes);
#ifdef SG_CHECK_QUEUES
        const size_t chk_bytes = (8 + 4 * (size_t)n_envs) * sizeof(uint32_t);
        if (err == cudaSuccess) err = cudaMalloc((void**)&h->d_chk, chk_bytes);
        if (err == cudaSuccess) err = cudaMemset(h->d_chk, 0, chk_bytes);
#endif
        h->q_bytes_total = q_bytes;
        for (int s = 0; s < h->n_sub && err == cudaSuccess && h->n_sub > 1; ++s) {
            err = cudaStreamCreateWithFlags(&h->sub[s].stream, cudaStreamNonBlocking);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h->sub[s].done, cudaEventDisableTiming);
        }
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming);
        const char* env_fused = getenv("SG_FUSED");
        h->use_fused = env_fused ? atoi(env_fused) : 0;
        const char* env_sg = getenv("SG_STAGED");
        h->use_staged = env_sg ? atoi(env_sg) : 0;  // measured neutral (profiles/r1_notes.md)
        const char* env_c1 = getenv("SG_CONT1");
        h->use_cont1 = env_c1 ? atoi(env_c1) : 0;
        const char* env_tm = getenv("SG_TIMING");
        h->timing = env_tm ? atoi(env_tm) : 0;
        h->t_n = 0;
        for (int k = 0; k < 4; ++k) h->t_acc[k] = 0;
        for (int k = 0; k < 3; ++k) h->q_acc[k] = 0;
        for (int k = 0; k < 5; ++k) {
            h->tev[k] = nullptr;
            if (h->timing && err == cudaSuccess) err = cudaEventCreate(&h->tev[k]);
        }
        const char* env_pdl = getenv("SG_PDL");
        h->use_pdl = env_pdl ? atoi(env_pdl) : 1;
        const char* env_tail = getenv("SG_TAIL");
        h->tail_mode = env_tail ? atoi(env_tail) : 2;
        const char* env_tg = getenv("SG_TAIL_CTAS");
        h->tail_grid = (env_tg && atoi(env_tg) > 0) ? atoi(env_tg) : 1;  // measured 1 / 2 / 4: 5.72 / 5.67 / 5.43e9
        const char* env_pri = getenv("SG_PRIORITY");
        h->use_priority = env_pri ? atoi(env_pri) : 0;
        if (err == cudaSuccess) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi is the numerically lowest = greatest priority
            err = cudaStreamCreateWithPriority(&h->hi_stream, cudaStreamNonBlocking, hi);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h->ev_hot, cudaEventDisableTiming);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h->ev_tail, cudaEventDisableTiming);
        }
    }
    if (err == cudaSuccess) err = cudaMemset(h->state, 0, (size_t)h->lay.total_bytes);
    if (err == cudaSuccess) err = cudaMemset(h->d_stats, 0, 4 * sizeof(double));
    if (err != cudaSuccess) {
        std::string msg = std::string("sg_create: ") + cudaGetErrorString(err);
        delete h;
        return fail(nullptr, SG_ERR_CUDA, msg);
    }
    build_dev(h);
    for (int s = 0; s < h->n_sub; ++s)  // continuation index words start out "empty" (fused kernel protocol)
        cudaMemset(h->sub[s].dev.q_cont, 0xFF, (size_t)h->sub[s].n * sizeof(uint32_t));
    if (c.task == SG_TASK_KEPLER) {
        int blocks = (int)((n_envs + 255) / 256);
        fill_orbit_kernel<<<blocks, 256>>>(h->dev.orbit, n_envs, c.ref_orbit_eccentricity, c.ref_orbit_angle);
        h->launches += 1;
    }
    err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        std::string msg = std::string("sg_create: ") + cudaGetErrorString(err);
        delete h;
        return fail(nullptr, SG_ERR_CUDA, msg);
    }
    *out = h;
    return SG_OK;
}

extern "C" void sg_destroy(sg_env* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->owns_state && h->state) cudaFree(h->state);
    if (h->d_stats) cudaFree(h->d_stats);
    if (h->d_queues) cudaFree(h->d_queues);
    if (h->d_chk) cudaFree(h->d_chk);
    for (int g = 0; g < h->n_graphs; ++g)
        if (h->graphs[g].exec) cudaGraphExecDestroy(h->graphs[g].exec);
    for (int s = 0; s < 8; ++s) {
        if (h->sub[s].stream) cudaStreamDestroy(h->sub[s].stream);
        if (h->sub[s].done) cudaEventDestroy(h->sub[s].done);
    }
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->hi_stream) cudaStreamDestroy(h->hi_stream);
    if (h->ev_hot) cudaEventDestroy(h->ev_hot);
    if (h->ev_tail) cudaEventDestroy(h->ev_tail);
    if (h->s_actions) cudaFree(h->s_actions);
    if (h->s_block) cudaFree(h->s_block);
    if (h->h_after) cudaEventDestroy(h->h_after);
    if (h->s_reward) cudaFree(h->s_reward);
    if (h->s_done) cudaFree(h->s_done);
    if (h->s_trunc) cudaFree(h->s_trunc);
    if (h->hstream) cudaStreamDestroy(h->hstream);
    delete h;
}

extern "C" int sg_get_layout(const sg_env* h, sg_layout* out) {
    if (!h || !out) return SG_ERR_INVALID;
    *out = h->lay;
    return SG_OK;
}
extern "C" void* sg_state_ptr(const sg_env* h) { return h ? h->state : nullptr; }
extern "C" uint64_t sg_step_count(const sg_env* h) { return h ? h->step_count : 0; }
static int sync_step_counters(sg_env* h, cudaStream_t stream) {
    // one launch for all sub-batches: thread s writes the counter of sub-batch s
    StepCounterPtrs p{};
    for (int s = 0; s < h->n_sub; ++s) p.q[s] = h->sub[s].dev.q_counts;
    set_step_counter_kernel<<<1, 8, 0, stream>>>(p, h->n_sub, h->step_count);
    h->launches += 1;
    return SG_OK;
}

extern "C" int sg_set_step_count(sg_env* h, uint64_t s) {
    if (!h) return SG_ERR_INVALID;
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaDeviceSynchronize());
    h->step_count = s;
    sync_step_counters(h, nullptr);
    SG_CUDA(h, cudaDeviceSynchronize());
    return SG_OK;
}
extern "C" int64_t sg_kernel_launches(const sg_env* h) { return h ? h->launches : 0; }

// env.seed(seed) (spaceship_env.py:92-94, goal.py:74-77): re-key the per-env Philox streams.  Every env's draw
// counter and the step counter that keys the synthetic-action stream restart at 0, so seed(s); reset(); ... replays
// exactly what a handle created with seed s does.  Cached graphs hold the old key in their kernel parameters.
extern "C" int sg_reseed(sg_env* h, uint64_t seed) {
    if (!h) return SG_ERR_INVALID;
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaDeviceSynchronize());
    h->seed = seed;
    build_dev(h);
    for (int g = 0; g < h->n_graphs; ++g)
        if (h->graphs[g].exec) cudaGraphExecDestroy(h->graphs[g].exec);
    h->n_graphs = 0;
    h->graph_misses = 0;
    memset(h->graphs, 0, sizeof(h->graphs));
    SG_CUDA(h, cudaMemset(h->dev.rng_ctr, 0, (size_t)h->n * sizeof(uint32_t)));
    h->step_count = 0;
    sync_step_counters(h, nullptr);
    SG_CUDA(h, cudaDeviceSynchronize());
    return SG_OK;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// Terminal observations under auto-reset (gym.vector-style trainers bootstrap from them when an episode is truncated by
// the TimeLimit): rows [N][D] of the dtype of the step's d_obs.  Every step call that writes an observation also writes,
// for each env that restarted in that step, the observation of the state its episode ended in; rows of other envs are
// left as they are.  NULL turns it off.  Written by the tail kernel only (terminal events: computed there; TimeLimit:
// the row the hot / continuation kernel wrote, saved before the reset observation replaces it).
extern "C" int sg_set_terminal_obs(sg_env* h, void* d_term_obs) {
    if (!h) return SG_ERR_INVALID;
    if (d_term_obs && !aligned16(d_term_obs)) return fail(h, SG_ERR_INVALID, "sg_set_terminal_obs: buffer must be 16-byte aligned");
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaDeviceSynchronize());
    h->term_obs = d_term_obs;
    for (int g = 0; g < h->n_graphs; ++g)  // cached graphs hold the old pointer in their kernel parameters
        if (h->graphs[g].exec) cudaGraphExecDestroy(h->graphs[g].exec);
    h->n_graphs = 0;
    h->graph_misses = 0;
    memset(h->graphs, 0, sizeof(h->graphs));
    return SG_OK;
}

// -DSG_CHECK_QUEUES builds: the device-side invariant counters of the split step's work queues (DevCfg::chk_err),
// out[0..5] = violations by kind, out[6] = env steps checked.  Other builds: SG_ERR_INVALID ("not compiled in").
extern "C" int sg_debug_queue_check(sg_env* h, uint64_t out[8]) {
    if (!h || !out) return SG_ERR_INVALID;
#ifdef SG_CHECK_QUEUES
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaDeviceSynchronize());
    uint32_t e[8];
    SG_CUDA(h, cudaMemcpy(e, h->d_chk, sizeof(e), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out[k] = e[k];
    out[6] = h->chk_steps;
    out[7] = 1;
    return SG_OK;
#else
    for (int k = 0; k < 8; ++k) out[k] = 0;
    return fail(h, SG_ERR_INVALID, "sg_debug_queue_check: this build was compiled without -DSG_CHECK_QUEUES");
#endif
}

// debugging aid (SG_TIMING=1): mean milliseconds of {hot, continuation, tail} per env step so far
extern "C" int sg_debug_kernel_times(sg_env* h, double out[6], int reset) {
    if (!h || !out) return SG_ERR_INVALID;
    for (int k = 0; k < 3; ++k) out[k] = h->t_n ? h->t_acc[k] / (double)h->t_n : 0.0;
    for (int k = 0; k < 3; ++k) out[3 + k] = h->t_n ? h->q_acc[k] / (double)h->t_n : 0.0;
    if (reset) {
        h->t_n = 0;
        for (int k = 0; k < 4; ++k) h->t_acc[k] = 0;
        for (int k = 0; k < 3; ++k) h->q_acc[k] = 0;
    }
    return SG_OK;
}

#endif  // SG_MAIN_PART

// ---- kernel dispatch over (task, planets, steering, action space, obs dtype) -------------------
enum Op { OP_STEP, OP_STEP_SPLIT, OP_RESET, OP_VFIELD };

struct LaunchArgs {
    Op op;
    cudaStream_t stream;
    const float* actions;
    int K;
    unsigned long long step0;
    void* obs;
    double* reward;
    uint8_t* done;
    uint8_t* trunc;
    const uint8_t* mask;
    double* vf_out;
    int obs_f32;     // dtype of `obs` for this call (the host result block may differ from the handle's obs_dtype)
    int pack_flags;  // done | truncated << 1 in one byte (host result block)
    int sub;  // OP_STEP_SPLIT: sub-batch index
    cudaStream_t hi_stream;  // OP_STEP_SPLIT: stream of the small kernels (nullptr: same stream)
};

template <typename... KArgs, typename... Args>
static void launch_k_smem(void (*kernel)(KArgs...), int grid, int block, int dyn_smem, cudaStream_t stream, bool pdl,
                          Args... args);
// Launch with (pdl = true) programmatic stream serialization: the grid may start while the previous kernel of the
// stream drains; the kernels order themselves with pdl_wait().
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), int grid, int block, cudaStream_t stream, bool pdl, Args... args) {
    launch_k_smem(kernel, grid, block, 0, stream, pdl, args...);
}
template <typename... KArgs, typename... Args>
static void launch_k_smem(void (*kernel)(KArgs...), int grid, int block, int dyn_smem, cudaStream_t stream, bool pdl,
                          Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = (size_t)dyn_smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <class TR, typename OT>
static void launch_typed(const sg_env* h, const LaunchArgs& a) {
    const int blocks = (int)((h->n + kBlock - 1) / kBlock);
    if (a.op == OP_STEP) {
        DevCfg dv = h->dev;
        dv.pack_flags = a.pack_flags;
        step_kernel<TR, OT><<<blocks, kBlock, 0, a.stream>>>(dv, a.actions, a.K, a.step0, (OT*)a.obs, a.reward, a.done,
                                                             a.trunc);
    } else if (a.op == OP_STEP_SPLIT) {
        const sg_env::Sub& sb = h->sub[a.sub];
        DevCfg dv = sb.dev;  // kernel parameter of this launch
        dv.pack_flags = a.pack_flags;
        // terminal rows have the handle's observation dtype: device-path calls only (the host result block brings its
        // own dtype and packed flags)
        const bool term_ok = h->term_obs && a.obs && h->tail_mode != 1 && !a.pack_flags &&
                             (a.obs_f32 != 0) == (h->cfg.obs_dtype == SG_OBS_F32);
        dv.term_obs = term_ok ? (void*)((char*)h->term_obs + (size_t)sb.off * TR::D * sizeof(OT)) : nullptr;
        const long long off = sb.off;
        const int sms = h->sm_count > 0 ? h->sm_count : 148;
        const int hot_blocks = (int)((sb.n + kBlock - 1) / kBlock);
        const int cont_blocks = hot_blocks < sms * SG_CONT_GRID ? hot_blocks : sms * SG_CONT_GRID;
        const long long tb = (sb.n + kTailBlock - 1) / kTailBlock;
        const int tail_blocks = (int)(tb < sms * SG_TAIL_GRID ? tb : sms * SG_TAIL_GRID);
        const float* act = a.actions ? a.actions + 2 * off : nullptr;
        OT* obs = a.obs ? (OT*)a.obs + off * TR::D : nullptr;
        double* rew = a.reward ? a.reward + off : nullptr;
        uint8_t* dn = a.done ? a.done + off : nullptr;
        uint8_t* tr = a.trunc ? a.trunc + off : nullptr;
        sg_env* hm = const_cast<sg_env*>(h);
#ifdef SG_CHECK_QUEUES
        dv.chk_epoch = h->chk_steps + 1;
        chk_pre_kernel<<<1, 32, 0, a.stream>>>(dv);
#endif
        if (h->timing) cudaEventRecord(h->tev[0], a.stream);
        if (h->use_fused) {
            int fused_blocks = hot_blocks < sms * SG_HOT_MINB ? hot_blocks : sms * SG_HOT_MINB;
            fused_kernel<TR, OT><<<fused_blocks, kBlock, 0, a.stream>>>(dv, act, obs, rew, dn, tr, (uint32_t)hot_blocks);
        } else if (h->use_staged) {
            int pblocks = hot_blocks < sms * SG_HOT_MINB ? hot_blocks : sms * SG_HOT_MINB;
            static bool attr_set_dev[64] = {};  // function attributes are per device (and per instantiation)
            bool& attr_set = attr_set_dev[h->device & 63];
            if (!attr_set) {
                cudaFuncSetAttribute(hot_staged_kernel<TR, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(HotPipeSmem<TR, OT>));
                cudaFuncSetAttribute(hot_staged_kernel<TR, OT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared);
                attr_set = true;
            }
            hot_staged_kernel<TR, OT><<<pblocks, kBlock, sizeof(HotPipeSmem<TR, OT>), a.stream>>>(dv, act, obs, rew, dn, tr,
                                                                                             (uint32_t)hot_blocks);
        } else {
            // SG_HOT_DYNSMEM=<bytes> (measurement aid): unused dynamic shared memory that limits the hot kernel to fewer
            // CTAs per SM with the SAME binary (38912 -> 3 CTAs, 97280 -> 2), to read its sensitivity to occupancy
            static const int hot_dyn = getenv("SG_HOT_DYNSMEM") ? atoi(getenv("SG_HOT_DYNSMEM")) : 0;
            if (hot_dyn > 0)
                cudaFuncSetAttribute(hot_kernel<TR, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, hot_dyn);
            launch_k_smem(hot_kernel<TR, OT>, hot_blocks, kBlock, hot_dyn, a.stream, h->use_pdl != 0, dv, act, obs, rew, dn, tr);
        }
        if (h->timing) cudaEventRecord(h->tev[1], a.stream);
        cudaStream_t small = a.stream;
        if (a.hi_stream) {
            small = a.hi_stream;
            cudaEventRecord(h->ev_hot, a.stream);
            cudaStreamWaitEvent(small, h->ev_hot, 0);
        }
        // step_plan 0 (default): hot -> generic continuation loop -> tail.  1: hot -> single-attempt continuation
        // -> loop for the rest -> tail.  2: fused persistent hot+continuation -> loop for the rest -> tail.
        if (h->use_fused == 0 && h->use_cont1) cont1_kernel<TR, OT><<<cont_blocks, kBlock, 0, small>>>(dv, act, obs, rew, dn, tr);
        if (h->use_fused || h->use_cont1)
            cont_kernel<TR, OT><<<sms, kBlock, 0, small>>>(dv, act, obs, rew, dn, tr, 1);
        else
            launch_k(cont_kernel<TR, OT>, cont_blocks, kBlock, small, h->use_pdl != 0 && !a.hi_stream, dv, act, obs, rew,
                     dn, tr, 0);
        if (h->timing) {
            cudaEventRecord(h->tev[2], small);
            uint32_t qc[3] = {0, 0, 0};
            cudaMemcpyAsync(qc, sb.dev.q_counts, sizeof(qc), cudaMemcpyDeviceToHost, small);
            cudaStreamSynchronize(small);
            for (int k = 0; k < 3; ++k) hm->q_acc[k] += qc[k];
            cudaEventRecord(h->tev[4], small);
        }
        if (h->tail_mode == 1) {
            tail_kernel<TR, OT><<<tail_blocks, kTailBlock, 0, small>>>(dv, act, obs, rew, dn, tr);
        } else {
            const long long t2 = (sb.n + 31) / 32;
            const int tail2_blocks = (int)(t2 < (long long)sms * h->tail_grid ? t2 : (long long)sms * h->tail_grid);
            launch_k(tail2_kernel<TR, OT>, tail2_blocks, kTail2Block, small, h->use_pdl != 0 && !h->timing, dv, act, obs,
                     rew, dn, tr);
        }
#ifdef SG_CHECK_QUEUES
        chk_post_kernel<<<sms, 256, 0, small>>>(dv, h->chk_steps + 1);
#endif
        if (h->timing) {  // debugging aid: synchronous, never on in benchmarks
            cudaEventRecord(h->tev[3], small);
            cudaEventSynchronize(h->tev[3]);
            float ms = 0;
            cudaEventElapsedTime(&ms, h->tev[0], h->tev[1]);
            hm->t_acc[0] += ms;
            cudaEventElapsedTime(&ms, h->tev[1], h->tev[2]);
            hm->t_acc[1] += ms;
            cudaEventElapsedTime(&ms, h->tev[4], h->tev[3]);
            hm->t_acc[2] += ms;
            hm->t_n += 1;
        }
        if (a.hi_stream) {
            cudaEventRecord(h->ev_tail, small);
            cudaStreamWaitEvent(a.stream, h->ev_tail, 0);
        }
    } else if (a.op == OP_RESET) {
        reset_kernel<TR, OT><<<blocks, kBlock, 0, a.stream>>>(h->dev, a.mask, (OT*)a.obs);
    }
}

template <class TR>
static void launch_tr(const sg_env* h, const LaunchArgs& a) {
    if (a.op == OP_VFIELD) {
        const int blocks = (int)((h->n + kBlock - 1) / kBlock);
        vector_field_kernel<TR><<<blocks, kBlock, 0, a.stream>>>(h->dev, a.actions, a.vf_out);
        return;
    }
    if (a.obs_f32) launch_typed<TR, float>(h, a);
    else launch_typed<TR, double>(h, a);
}

template <int TASK, int NP>
void launch_np(const sg_env* h, const LaunchArgs& a) {
    const bool vel = h->cfg.ship_steering == SG_STEER_VELOCITY;
    const bool disc = h->cfg.discrete != 0;
    if (vel && !disc) launch_tr<Traits<TASK, NP, STEER_VEL, 0>>(h, a);
    else if (vel && disc) launch_tr<Traits<TASK, NP, STEER_VEL, 1>>(h, a);
    else if (!vel && !disc) launch_tr<Traits<TASK, NP, STEER_ACC, 0>>(h, a);
    else launch_tr<Traits<TASK, NP, STEER_ACC, 1>>(h, a);
}

#if defined(SG_PART)
#if SG_PART == 0  // the other families are instantiated by parts 1-5
extern template void launch_np<TASK_GOAL, 1>(const sg_env*, const LaunchArgs&);
extern template void launch_np<TASK_GOAL, 3>(const sg_env*, const LaunchArgs&);
extern template void launch_np<TASK_GOAL, 4>(const sg_env*, const LaunchArgs&);
extern template void launch_np<TASK_KEPLER, 1>(const sg_env*, const LaunchArgs&);
extern template void launch_np<TASK_DNC, 1>(const sg_env*, const LaunchArgs&);
#elif SG_PART == 1
template void launch_np<TASK_GOAL, 3>(const sg_env*, const LaunchArgs&);
#elif SG_PART == 2
template void launch_np<TASK_GOAL, 4>(const sg_env*, const LaunchArgs&);
#elif SG_PART == 3
template void launch_np<TASK_KEPLER, 1>(const sg_env*, const LaunchArgs&);
#elif SG_PART == 4
template void launch_np<TASK_DNC, 1>(const sg_env*, const LaunchArgs&);
#elif SG_PART == 5
template void launch_np<TASK_GOAL, 1>(const sg_env*, const LaunchArgs&);
#endif
#endif

#if SG_MAIN_PART
// enqueue only (no error query: also used inside stream capture)
static int enqueue(sg_env* h, const LaunchArgs& a) {
#ifdef SG_DEV_ONLY_GOAL3P  // development builds (seconds instead of minutes): the BASELINE workload's instantiation only
    if (h->cfg.task == SG_TASK_GOAL && h->cfg.n_planets == 3 && h->cfg.ship_steering == SG_STEER_VELOCITY &&
        !h->cfg.discrete && a.op != OP_VFIELD) {
        if (a.obs_f32) launch_typed<Traits<TASK_GOAL, 3, STEER_VEL, 0>, float>(h, a);
        else launch_typed<Traits<TASK_GOAL, 3, STEER_VEL, 0>, double>(h, a);
    }
    else return fail(h, SG_ERR_INVALID, "SG_DEV_ONLY_GOAL3P build: only GoalContinuous3P-v0");
#else
    if (h->cfg.task == SG_TASK_KEPLER) launch_np<TASK_KEPLER, 1>(h, a);
    else if (h->cfg.task == SG_TASK_DONOTCRASH) launch_np<TASK_DNC, 1>(h, a);
    else if (h->cfg.n_planets == 1) launch_np<TASK_GOAL, 1>(h, a);
    else if (h->cfg.n_planets == 2) launch_np<TASK_GOAL, 2>(h, a);
    else if (h->cfg.n_planets == 3) launch_np<TASK_GOAL, 3>(h, a);
    else if (h->cfg.n_planets == 4) launch_np<TASK_GOAL, 4>(h, a);
    else return fail(h, SG_ERR_INVALID, "unsupported n_planets");
#endif
    h->launches += (a.op == OP_STEP_SPLIT) ? ((h->use_cont1 && !h->use_fused) ? 4 : 3) : 1;
    return SG_OK;
}

static int launch(sg_env* h, const LaunchArgs& a) {
    SG_CUDA(h, cudaSetDevice(h->device));
    int rc = enqueue(h, a);
    if (rc != SG_OK) return rc;
    SG_CUDA(h, cudaGetLastError());
    return SG_OK;
}

// One env step of the split path for every sub-batch: fork from `origin`, three kernels per sub-batch on its
// own stream, join back into `origin`.
static int enqueue_split_step(sg_env* h, cudaStream_t origin, LaunchArgs a) {
    a.op = OP_STEP_SPLIT;
    if (h->n_sub == 1) {
        a.sub = 0;
        a.stream = origin;
        a.hi_stream = (h->use_priority == 1 && origin != h->hi_stream) ? h->hi_stream : nullptr;
        return enqueue(h, a);
    }
    a.hi_stream = nullptr;
    SG_CUDA(h, cudaEventRecord(h->fork_ev, origin));
    for (int s = 0; s < h->n_sub; ++s) {
        cudaStream_t st = h->sub[s].stream;
        if (st != origin) SG_CUDA(h, cudaStreamWaitEvent(st, h->fork_ev, 0));
        a.sub = s;
        a.stream = st;
        int rc = enqueue(h, a);
        if (rc != SG_OK) return rc;
    }
    for (int s = 0; s < h->n_sub; ++s) {
        cudaStream_t st = h->sub[s].stream;
        if (st == origin) continue;
        SG_CUDA(h, cudaEventRecord(h->sub[s].done, st));
        SG_CUDA(h, cudaStreamWaitEvent(origin, h->sub[s].done, 0));
    }
    return SG_OK;
}

// Same, replayed from a cached CUDA graph keyed by the argument pointers (one launch per env step).
static int launch_split_step(sg_env* h, cudaStream_t user, const LaunchArgs& a, bool may_graph) {
    SG_CUDA(h, cudaSetDevice(h->device));
    // Graph replay only pays for the fork / join over several sub-batches, and only when the same argument pointers
    // come back (single-step calls on persistent buffers).  The K steps of a multi-step call each have their own
    // reward / done offsets (K distinct keys: every step would capture and instantiate), and a caller that passes
    // fresh tensors every step never hits: after 32 consecutive misses the handle stops capturing.
    if (!h->use_graphs || h->n_sub == 1 || !may_graph || h->graph_misses > 32) {
        int rc = enqueue_split_step(h, user, a);
        if (rc != SG_OK) return rc;
        SG_CUDA(h, cudaGetLastError());
        return SG_OK;
    }
    const void* key[6] = {a.actions, a.obs, a.reward, a.done, a.trunc,
                          (const void*)(uintptr_t)(1u | (a.obs_f32 ? 2u : 0u) | (a.pack_flags ? 4u : 0u))};
    sg_env::GraphEntry* hit = nullptr;
    for (int g = 0; g < h->n_graphs; ++g)
        if (!memcmp(h->graphs[g].key, key, sizeof(key))) hit = &h->graphs[g];
    if (hit) h->graph_misses = 0;
    if (!hit) {
        h->graph_misses += 1;
        // capture is not allowed on the legacy default stream: capture on an internal one
        cudaStream_t origin = h->n_sub > 1 ? h->sub[0].stream : h->hi_stream;
        cudaGraph_t graph = nullptr;
        const int64_t launches_before = h->launches;
        SG_CUDA(h, cudaStreamBeginCapture(origin, cudaStreamCaptureModeRelaxed));
        int rc = enqueue_split_step(h, origin, a);
        const cudaError_t lerr = cudaGetLastError();  // a launch that failed inside the capture invalidates it
        cudaError_t cerr = cudaStreamEndCapture(origin, &graph);
        h->launches = launches_before;
        if (rc != SG_OK) return rc;
        if (lerr != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return fail(h, SG_ERR_CUDA, std::string("kernel launch during graph capture: ") + cudaGetErrorString(lerr));
        }
        if (cerr != cudaSuccess) return fail(h, SG_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(cerr));
        int slot = h->n_graphs;
        if (slot == 16) {  // evict the least recently used entry
            slot = 0;
            for (int g = 1; g < 16; ++g)
                if (h->graphs[g].stamp < h->graphs[slot].stamp) slot = g;
            cudaGraphExecDestroy(h->graphs[slot].exec);
        } else {
            h->n_graphs += 1;
        }
        hit = &h->graphs[slot];
        memcpy(hit->key, key, sizeof(key));
        cerr = cudaGraphInstantiate(&hit->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (cerr != cudaSuccess) {
            hit->exec = nullptr;
            memset(hit->key, 0xff, sizeof(hit->key));
            return fail(h, SG_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(cerr));
        }
    }
    hit->stamp = ++h->graph_clock;
    SG_CUDA(h, cudaGraphLaunch(hit->exec, user));
    h->launches += 3 * h->n_sub;
    return SG_OK;
}


extern "C" int sg_reset(sg_env* h, const uint8_t* d_mask, void* d_obs, void* stream) {
    if (!h) return SG_ERR_INVALID;
    LaunchArgs a{};
    a.op = OP_RESET;
    a.stream = (cudaStream_t)stream;
    a.mask = d_mask;
    a.obs = d_obs;
    a.obs_f32 = h->cfg.obs_dtype == SG_OBS_F32;
    return launch(h, a);
}

// K env steps; obs_f32 / pack_flags describe the output buffers of THIS call (the host result block asks for
// float32 observations and packed flags whatever the handle's obs_dtype is)
static int step_impl(sg_env* h, const float* d_actions, int K, void* d_obs, int obs_f32, double* d_reward, uint8_t* d_done,
                     uint8_t* d_truncated, int pack_flags, cudaStream_t stream) {
    if (!h || K <= 0) return SG_ERR_INVALID;
    if (d_obs && !aligned16(d_obs)) return fail(h, SG_ERR_INVALID, "sg_step: d_obs must be 16-byte aligned");
    if (d_actions && (((uintptr_t)d_actions) & 7u)) return fail(h, SG_ERR_INVALID, "sg_step: d_actions must be 8-byte aligned");
    const int mode = h->cfg.step_mode;
    // auto = split for every K: K split steps (3 K launches) run 3-4 x faster than the register-resident K-fused
    // kernel (3.3 vs 1.05 G env-steps/s on one 262144-env handle), whose warps pay for every rare path
    const bool split = (mode == SG_STEP_SPLIT) || (mode == SG_STEP_AUTO);
    LaunchArgs a{};
    a.stream = stream;
    a.obs_f32 = obs_f32;
    a.pack_flags = pack_flags;
    if (!split) {
        a.op = OP_STEP;
        a.actions = d_actions;
        a.K = K;
        a.step0 = h->step_count;
        a.obs = d_obs;
        a.reward = d_reward;
        a.done = d_done;
        a.trunc = d_truncated;
        int rc = launch(h, a);
        if (rc != SG_OK) return rc;
        h->step_count += (uint64_t)K;
        sync_step_counters(h, a.stream);  // keep the device-side counters of the split path in step
        SG_CUDA(h, cudaGetLastError());
        return SG_OK;
    }
    // split path: K single steps (hot / continuation / tail kernels per sub-batch); the observation is only
    // needed after the last one
    a.op = OP_STEP_SPLIT;
    a.K = 1;
    for (int k = 0; k < K; ++k) {
        const size_t off = (size_t)k * (size_t)h->n;
        a.actions = d_actions ? d_actions + 2 * off : nullptr;
        a.obs = (k == K - 1) ? d_obs : nullptr;
        a.reward = d_reward ? d_reward + off : nullptr;
        a.done = d_done ? d_done + off : nullptr;
        a.trunc = d_truncated ? d_truncated + off : nullptr;
        int rc = launch_split_step(h, stream, a, K == 1);
        if (rc != SG_OK) return rc;
        h->step_count += 1;  // the tail kernels advance the device-side counters
        h->chk_steps += 1;
    }
    return SG_OK;
}

extern "C" int sg_step(sg_env* h, const float* d_actions, int K, void* d_obs, double* d_reward, uint8_t* d_done,
                       uint8_t* d_truncated, void* stream) {
    if (!h) return SG_ERR_INVALID;
    return step_impl(h, d_actions, K, d_obs, h->cfg.obs_dtype == SG_OBS_F32, d_reward, d_done, d_truncated, 0,
                     (cudaStream_t)stream);
}

extern "C" int sg_vector_field(sg_env* h, const float* d_actions, double* d_out, void* stream) {
    if (!h || !d_actions || !d_out) return SG_ERR_INVALID;
    LaunchArgs a{};
    a.op = OP_VFIELD;
    a.stream = (cudaStream_t)stream;
    a.actions = d_actions;
    a.vf_out = d_out;
    return launch(h, a);
}

// ---- host-buffer entry points --------------------------------------------------------------------------------
// One device-side result block per handle, laid out like the host block the caller passes (sg_host_layout):
//   obs [N][D] (f32 or f64) | reward f64 [N] | flags u8 [N] (bit 0 done, bit 1 truncated), 256-byte aligned parts.
// A step is ONE host->device copy (actions), the step kernels, ONE device->host copy (the block), all on the
// handle's own non-blocking stream.  That stream is ordered after `after_stream` (the stream the caller's other
// work on this handle runs on: resets, set_state, device-path steps) by an event, so reset(); step(numpy) is safe.
static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" int sg_host_block_layout(const sg_env* h, int obs_dtype, int want_obs, sg_host_layout* out) {
    if (!h || !out) return SG_ERR_INVALID;
    const int64_t n = h->n, D = sg_obs_dim(&h->cfg);
    memset(out, 0, sizeof(*out));
    out->obs_dtype = obs_dtype;
    out->obs_dim = (int32_t)D;
    out->n_envs = n;
    int64_t off = 0;
    out->obs = off;
    off += want_obs ? align256(n * D * (obs_dtype == SG_OBS_F32 ? 4 : 8)) : 0;
    out->reward = off;
    off += align256(n * 8);
    out->flags = off;
    off += align256(n);
    out->total_bytes = off;
    return SG_OK;
}

static int ensure_host_path(sg_env* h, int K) {
    SG_CUDA(h, cudaSetDevice(h->device));
    if (!h->hstream) {
        SG_CUDA(h, cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
        SG_CUDA(h, cudaEventCreateWithFlags(&h->h_after, cudaEventDisableTiming));
    }
    const size_t n = (size_t)h->n, D = (size_t)sg_obs_dim(&h->cfg);
    if (!h->s_block) {  // sized for the largest layout (f64 observations)
        h->s_block_bytes = (size_t)(align256((int64_t)(n * D * 8)) + align256((int64_t)(n * 8)) + align256((int64_t)n));
        SG_CUDA(h, cudaMalloc(&h->s_block, h->s_block_bytes));
    }
    if ((size_t)K > h->s_k) {  // legacy multi-step form: per-step reward / done / truncated rows
        if (h->s_actions) cudaFree(h->s_actions);
        if (h->s_reward) cudaFree(h->s_reward);
        if (h->s_done) cudaFree(h->s_done);
        if (h->s_trunc) cudaFree(h->s_trunc);
        h->s_reward = nullptr;
        h->s_done = h->s_trunc = nullptr;
        SG_CUDA(h, cudaMalloc((void**)&h->s_actions, (size_t)K * n * 2 * sizeof(float)));
        if (K > 1) {
            SG_CUDA(h, cudaMalloc((void**)&h->s_reward, (size_t)K * n * sizeof(double)));
            SG_CUDA(h, cudaMalloc((void**)&h->s_done, (size_t)K * n));
            SG_CUDA(h, cudaMalloc((void**)&h->s_trunc, (size_t)K * n));
        }
        h->s_k = (size_t)K;
    }
    return SG_OK;
}

// the handle's stream starts after everything enqueued so far on `after` (NULL: the legacy default stream, which
// itself follows every blocking stream)
static int order_after(sg_env* h, cudaStream_t after) {
    SG_CUDA(h, cudaEventRecord(h->h_after, after));
    SG_CUDA(h, cudaStreamWaitEvent(h->hstream, h->h_after, 0));
    return SG_OK;
}

extern "C" int sg_step_host_block_async(sg_env* h, const float* h_actions, void* h_block, int obs_dtype, int want_obs,
                                        void* after_stream) {
    if (!h || !h_block) return SG_ERR_INVALID;
    if (obs_dtype != SG_OBS_F32 && obs_dtype != SG_OBS_F64) return fail(h, SG_ERR_INVALID, "sg_step_host_block_async: bad obs_dtype");
    int rc = ensure_host_path(h, 1);
    if (rc != SG_OK) return rc;
    rc = order_after(h, (cudaStream_t)after_stream);
    if (rc != SG_OK) return rc;
    sg_host_layout lay;
    sg_host_block_layout(h, obs_dtype, want_obs, &lay);
    const size_t n = (size_t)h->n;
    char* blk = (char*)h->s_block;
    if (h_actions)
        SG_CUDA(h, cudaMemcpyAsync(h->s_actions, h_actions, n * 2 * sizeof(float), cudaMemcpyHostToDevice, h->hstream));
    rc = step_impl(h, h_actions ? h->s_actions : nullptr, 1, want_obs ? (void*)(blk + lay.obs) : nullptr,
                   obs_dtype == SG_OBS_F32, (double*)(blk + lay.reward), (uint8_t*)(blk + lay.flags), nullptr, 1, h->hstream);
    if (rc != SG_OK) return rc;
    SG_CUDA(h, cudaMemcpyAsync(h_block, blk, (size_t)lay.total_bytes, cudaMemcpyDeviceToHost, h->hstream));
    return SG_OK;
}

static size_t obs_elem(const sg_env* h) { return h->cfg.obs_dtype == SG_OBS_F32 ? 4 : 8; }

extern "C" int sg_reset_host(sg_env* h, void* h_obs) {
    if (!h) return SG_ERR_INVALID;
    int rc = ensure_host_path(h, 1);
    if (rc != SG_OK) return rc;
    rc = order_after(h, nullptr);
    if (rc != SG_OK) return rc;
    rc = sg_reset(h, nullptr, h_obs ? h->s_block : nullptr, h->hstream);
    if (rc != SG_OK) return rc;
    if (h_obs)
        SG_CUDA(h, cudaMemcpyAsync(h_obs, h->s_block, (size_t)h->n * sg_obs_dim(&h->cfg) * obs_elem(h),
                                   cudaMemcpyDeviceToHost, h->hstream));
    SG_CUDA(h, cudaStreamSynchronize(h->hstream));
    return SG_OK;
}

// separate-array form (K steps; observation in the handle's obs_dtype): five copies down per call
extern "C" int sg_step_host_async(sg_env* h, const float* h_actions, int K, void* h_obs, double* h_reward,
                                  uint8_t* h_done, uint8_t* h_truncated) {
    if (!h || K <= 0) return SG_ERR_INVALID;
    int rc = ensure_host_path(h, K > 1 ? K : 2);
    if (rc != SG_OK) return rc;
    rc = order_after(h, nullptr);
    if (rc != SG_OK) return rc;
    const size_t n = (size_t)h->n, kn = (size_t)K * n;
    if (h_actions)
        SG_CUDA(h, cudaMemcpyAsync(h->s_actions, h_actions, kn * 2 * sizeof(float), cudaMemcpyHostToDevice, h->hstream));
    rc = sg_step(h, h_actions ? h->s_actions : nullptr, K, h_obs ? h->s_block : nullptr, h_reward ? h->s_reward : nullptr,
                 h_done ? h->s_done : nullptr, h_truncated ? h->s_trunc : nullptr, h->hstream);
    if (rc != SG_OK) return rc;
    if (h_obs)
        SG_CUDA(h, cudaMemcpyAsync(h_obs, h->s_block, n * sg_obs_dim(&h->cfg) * obs_elem(h), cudaMemcpyDeviceToHost,
                                   h->hstream));
    if (h_reward) SG_CUDA(h, cudaMemcpyAsync(h_reward, h->s_reward, kn * sizeof(double), cudaMemcpyDeviceToHost, h->hstream));
    if (h_done) SG_CUDA(h, cudaMemcpyAsync(h_done, h->s_done, kn, cudaMemcpyDeviceToHost, h->hstream));
    if (h_truncated) SG_CUDA(h, cudaMemcpyAsync(h_truncated, h->s_trunc, kn, cudaMemcpyDeviceToHost, h->hstream));
    return SG_OK;
}

extern "C" int sg_host_wait(sg_env* h) {
    if (!h) return SG_ERR_INVALID;
    if (!h->hstream) return SG_OK;
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaStreamSynchronize(h->hstream));
    return SG_OK;
}

extern "C" int sg_step_host(sg_env* h, const float* h_actions, int K, void* h_obs, double* h_reward, uint8_t* h_done,
                            uint8_t* h_truncated) {
    int rc = sg_step_host_async(h, h_actions, K, h_obs, h_reward, h_done, h_truncated);
    if (rc != SG_OK) return rc;
    return sg_host_wait(h);
}

extern "C" int sg_stats(sg_env* h, double out[4], void* stream) {
    if (!h || !out) return SG_ERR_INVALID;
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaMemcpyAsync(out, h->d_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SG_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    return SG_OK;
}

extern "C" int sg_stats_reset(sg_env* h, void* stream) {
    if (!h) return SG_ERR_INVALID;
    SG_CUDA(h, cudaSetDevice(h->device));
    SG_CUDA(h, cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(double), (cudaStream_t)stream));
    return SG_OK;
}

extern "C" int sg_measure_fp64_peak(int device, double ms, double* tflops_out) {
    if (!tflops_out) return SG_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, SG_ERR_NO_DEVICE, "no CUDA device");
    SG_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    SG_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr;
    SG_CUDA(nullptr, cudaMalloc((void**)&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    SG_CUDA(nullptr, cudaEventCreate(&e0));
    SG_CUDA(nullptr, cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    int iters = 2000;
    double best = 0;
    dfma_peak_kernel<<<blocks, threads>>>(d_out, 200, 1.0);  // warm-up
    SG_CUDA(nullptr, cudaDeviceSynchronize());
    double spent = 0;
    for (int rep = 0; rep < 64 && spent < ms; ++rep) {
        SG_CUDA(nullptr, cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0 + rep);
        SG_CUDA(nullptr, cudaEventRecord(e1));
        SG_CUDA(nullptr, cudaEventSynchronize(e1));
        float t = 0;
        SG_CUDA(nullptr, cudaEventElapsedTime(&t, e0, e1));
        spent += t;
        double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        double tf = flops / (t * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops_out = best;
    return SG_OK;
}
#endif  // SG_MAIN_PART
