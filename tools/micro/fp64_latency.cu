// FP64 pipe of one B200 SM sub-partition: dependent-issue latency of DFMA / DMUL / DADD / MUFU.RSQ64H and the
// issue rate with W warps x ILP independent chains (nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu).
// One CTA on one SM; warps 0..W-1 of a 4*W-thread... layout: blockDim = 32*W*4 so that every scheduler holds W warps.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void rsq_chain(double* out, long long* cyc, int iters, double a) {
    double x = a + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y + 2.0;  // one DADD keeps the value in range: chain = MUFU.RSQ64H + DADD
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void dadd_chain(double* out, long long* cyc, int iters, double a) {
    double x = a + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x = x + a;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
static void run(int warps_per_sched, double* out, long long* cyc) {
    const int iters = 2000;
    const int threads = 32 * 4 * warps_per_sched;
    dfma_chain<ILP><<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    dfma_chain<ILP><<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16 * ILP;
    printf("{\"op\": \"DFMA\", \"warps_per_scheduler\": %d, \"ilp\": %d, \"cycles_per_dfma_per_warp\": %.3f, "
           "\"dfma_per_cycle_per_scheduler\": %.4f}\n",
           warps_per_sched, ILP, c / n, n * warps_per_sched / c);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1024);
    for (int w : {1, 2, 3, 4, 5, 6, 8}) {
        run<1>(w, out, cyc);
        run<2>(w, out, cyc);
        run<4>(w, out, cyc);
    }
    long long c = 0;
    rsq_chain<<<1, 32>>>(out, cyc, 2000, 3.0);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("{\"op\": \"MUFU.RSQ64H + DADD chain\", \"cycles_per_pair\": %.3f}\n", c / (2000.0 * 16));
    dadd_chain<<<1, 32>>>(out, cyc, 2000, 1.5);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("{\"op\": \"DADD chain\", \"cycles_per_dadd\": %.3f}\n", c / (2000.0 * 16));
    return 0;
}
