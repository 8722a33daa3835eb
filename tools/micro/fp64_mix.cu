// What the FP64 pipe of a B200 SM sustains on the hot kernel's OWN instruction mix (nvcc -arch=sm_100a -O3).
// One CTA per SM-sub-partition set: blockDim = 128 * W puts W warps on every scheduler.  Every kernel is a loop over
// register-resident work with no memory traffic; rates are per scheduler, 0.5 FP64 warp-instructions / cycle is the
// DFMA peak measured by fp64_latency.cu.
//   accel : the force evaluation of the engine (3 planets: rsqrt + cubic correction, inverse cube, accumulation),
//           each evaluation at a position that depends on the previous one (as the Runge-Kutta stages do)
//   accel2: two independent envs per thread interleaved (what a second warp would add, without the registers)
//   dmul / dadd / dsetp_sel / dfma_const / dfma_imad: single-opcode rates with 4 independent chains per warp
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double kC[8] = {0.3, 0.7, 1.1, 1.3, 1.7, 1.9, 2.3, 2.9};

__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}

__device__ __forceinline__ void accel3(const double (&px)[3], const double (&py)[3], double x, double y, double cth, double sth,
                                       double fs, double gm, double im, double& ax, double& ay) {
    double fx = -cth * fs, fy = -sth * fs;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        double dx = px[p] - x, dy = py[p] - y;
        double inv = fast_rsqrt(dx * dx + dy * dy);
        double s = gm * (inv * inv * inv);
        fx = fma(dx, s, fx);
        fy = fma(dy, s, fy);
    }
    ax = fx * im;
    ay = fy * im;
}

template <int ENVS>
__global__ void accel_kernel(double* out, long long* cyc, int iters, double seed) {
    double px[ENVS][3], py[ENVS][3], x[ENVS], y[ENVS];
#pragma unroll
    for (int e = 0; e < ENVS; ++e) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            px[e][p] = seed * (p + 1) + 0.001 * threadIdx.x + e;
            py[e][p] = -seed * (p + 2) + 0.002 * threadIdx.x - e;
        }
        x[e] = 0.01 * threadIdx.x;
        y[e] = 0.02 * threadIdx.x + e;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int e = 0; e < ENVS; ++e) {
                double ax, ay;
                accel3(px[e], py[e], x[e], y[e], 0.6, 0.8, 0.3, 1e-3, 0.5, ax, ay);
                x[e] = fma(1e-6, ax, x[e]);  // the next evaluation point depends on this force
                y[e] = fma(1e-6, ay, y[e]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int e = 0; e < ENVS; ++e) s += x[e] + y[e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void op_kernel(double* out, long long* cyc, int iters, double a, double b) {
    double x[4];
    int n[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = a + k + threadIdx.x;
        n[k] = threadIdx.x + k;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (OP == 0) x[k] = x[k] * b;                       // DMUL
                if (OP == 1) x[k] = x[k] + b;                       // DADD
                if (OP == 2) x[k] = (b > x[k]) ? b + x[k] * 0 : x[k];  // DSETP + selects (kept honest by the compiler or not: see SASS)
                if (OP == 3) x[k] = fma(x[k], kC[(r + k) & 7], kC[(r + 3) & 7]);  // DFMA with constant-bank operands
                if (OP == 4) {                                      // DFMA + an integer multiply-add per DFMA
                    x[k] = fma(x[k], b, a);
                    n[k] = n[k] * 3 + r;
                }
                if (OP == 5) x[k] = fma(x[k], b, a);               // DFMA (reference)
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += x[k] + n[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static long long get(long long* cyc) {
    long long c = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    return c;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 22);
    cudaMalloc(&cyc, 1024);
    const int iters = 500;
    // FP64-pipe instructions of one accel3 + update: counted from the SASS of this file (see the JSON field)
    for (int w : {1, 2, 3, 4, 5, 6, 8}) {
        for (int rep = 0; rep < 2; ++rep) accel_kernel<1><<<1, 128 * w>>>(out, cyc, iters, 1.37);
        long long c = get(cyc);
        printf("{\"kernel\": \"accel\", \"warps_per_scheduler\": %d, \"cycles_per_eval_per_warp\": %.2f, \"evals_per_kcycle_per_scheduler\": %.3f}\n",
               w, c / (iters * 4.0), 1000.0 * iters * 4.0 * w / c);
        if (w <= 4) {
            for (int rep = 0; rep < 2; ++rep) accel_kernel<2><<<1, 128 * w>>>(out, cyc, iters, 1.37);
            c = get(cyc);
            printf("{\"kernel\": \"accel2\", \"warps_per_scheduler\": %d, \"cycles_per_eval_per_warp\": %.2f, \"evals_per_kcycle_per_scheduler\": %.3f}\n",
                   w, c / (iters * 8.0), 1000.0 * iters * 8.0 * w / c);
        }
    }
    const char* names[6] = {"DMUL", "DADD", "DSETP+SEL", "DFMA const-bank", "DFMA + IMAD", "DFMA"};
    for (int op = 0; op < 6; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) op_kernel<0><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
            if (op == 1) op_kernel<1><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
            if (op == 2) op_kernel<2><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
            if (op == 3) op_kernel<3><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
            if (op == 4) op_kernel<4><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
            if (op == 5) op_kernel<5><<<1, 512>>>(out, cyc, 2000, 1.0000001, 0.9999999);
        }
        long long c = get(cyc);
        const double n = 2000.0 * 16 * 4;
        printf("{\"op\": \"%s\", \"warps_per_scheduler\": 4, \"ilp\": 4, \"per_cycle_per_scheduler\": %.4f}\n", names[op], n * 4 / c);
    }
    return 0;
}
