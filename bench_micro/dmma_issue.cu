// How fast can ONE warp issue fp64 DMMAs (mma.sync.m8n8k4.f64) on sm_100a, and what does it depend on?
//   fixed   16 independent accumulators, A/B operands loop-invariant registers
//   lds     A/B operands loaded from shared memory for every DMMA (conflict-free LDS.64), 16 independent accumulators
//   alu     A/B operands changed by an integer XOR before every DMMA (no loads)
//   chain   ONE accumulator: dependent DMMA latency
// Reported: cycles per DMMA per scheduler (SM sub-partition), for 1, 2, 4 warps per scheduler.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_issue dmma_issue.cu && ./dmma_issue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void k(double *out, int iters, long long *cycles)
{
    __shared__ double sm[32 * 64];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double c[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
    double a = 1.0 + lane * 1e-3, b = 1.0 - lane * 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma(c[j][0], c[j][1], a, b);
        } else if (MODE == 1) {
            const double *row = sm + ((it & 31) * 64) + lane;
            double f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = row[(j & 1) * 32];
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma(c[j][0], c[j][1], f[j], f[15 - j]);
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                a = __longlong_as_double(__double_as_longlong(a) ^ (long long)(it & 1));
                dmma(c[j][0], c[j][1], a, b);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma(c[0][0], c[0][1], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main()
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    long long *cyc, h = 0;
    cudaMalloc(&out, sizeof(double) * nsm * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 4000;
    const char *names[4] = {"fixed", "lds", "alu", "chain"};
    for (int mode = 0; mode < 4; ++mode)
        for (int wps = 1; wps <= 4; wps *= 2) {
            const int threads = 128 * wps;            // 4 schedulers x wps warps, one CTA per SM
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                case 0: k<0><<<nsm, threads>>>(out, iters, cyc); break;
                case 1: k<1><<<nsm, threads>>>(out, iters, cyc); break;
                case 2: k<2><<<nsm, threads>>>(out, iters, cyc); break;
                default: k<3><<<nsm, threads>>>(out, iters, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            printf("%-6s warps/scheduler=%d: %7.2f cycles per DMMA per scheduler (%7.2f per warp)\n", names[mode], wps,
                   (double)h / (16.0 * iters * wps), (double)h / (16.0 * iters));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
