// Microbenchmark: LATENCY of dependent fp64 / shuffle / MUFU chains on sm_100a (one warp, clock64), alone and with
// other warps of the same scheduler streaming DMMAs. Explains why the per-item tail of the BPMF kernel (a chain of
// dependent scalar fp64 instructions) is slow. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int N = 2048;

// mode: 0 DFMA chain, 1 DMUL->DFMA pairs, 2 64-bit SHFL chain, 3 MUFU.RCP64H + cubic chain, 4 SHFL + DFMA (one solve step),
//       5 two independent DFMA chains (ILP 2), 6 DMMA dependent chain (same accumulator)
// Warp 0 measures; warps 1.. (if any) stream independent DMMAs on the same SM until warp 0 is done.
__global__ void k_lat(int mode, long long *cycles, double *sink, volatile int *flag)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp > 0) {
        double c[8][2];
        for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
        double f = 1e-3 * lane;
        while (*flag == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], f, f);
        }
        double s = 0;
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        if (s == 123.456) sink[1] = s;
        return;
    }
    double x = 1.0 + 1e-9 * lane, y = 0.5 + 1e-9 * lane, a = 1.0000001, b = 1e-9;
    double c0 = 0, c1 = 0;
    long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = fma(x, a, b);
    } else if (mode == 1) {
#pragma unroll 16
        for (int i = 0; i < N / 2; ++i) { double t = x * a; x = fma(t, b, x); }
    } else if (mode == 2) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    } else if (mode == 3) {
#pragma unroll 4
        for (int i = 0; i < N / 4; ++i) {
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
            const double e = fma(-x, r, 1.0);
            const double t = fma(e, e, e);
            x = fma(r, t, r) + 1.5;
        }
    } else if (mode == 4) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) { const double t = __shfl_sync(0xffffffffu, x, i & 31); x = fma(-a, t, x); }
    } else if (mode == 5) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) { x = fma(x, a, b); y = fma(y, a, b); }
    } else {
#pragma unroll 16
        for (int i = 0; i < N; ++i) dmma884(c0, c1, x, y);
    }
    long long t1 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; *flag = 1; }
    if (x + y + c0 + c1 == 123.456) sink[0] = x;
}

int main()
{
    long long *cyc; double *sink; int *flag;
    CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&flag, 4));
    const char *names[] = {"DFMA -> DFMA", "DMUL -> DFMA (per op)", "SHFL.64 -> SHFL.64", "MUFU.RCP64H + 3 DFMA + DADD (per group)",
                           "SHFL.64 + DFMA (solve step)", "2 independent DFMA chains (per pair)", "DMMA -> DMMA same accumulator"};
    const int per[] = {N, N, N, N / 4, N, N, N};
    for (int busy : {0, 4, 8}) {          // extra DMMA-streaming warps in the block: 0, 1 per scheduler, 2 per scheduler
        printf("--- %d DMMA-streaming warps on the SM\n", busy);
        for (int mode = 0; mode < 7; ++mode) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(flag, 0, 4));
                k_lat<<<1, 32 * (1 + busy)>>>(mode, cyc, sink, flag);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            }
            printf("%-45s %7.1f cycles\n", names[mode], (double)h / per[mode]);
        }
    }
    return 0;
}
