// Microbenchmark: fp64 vector DFMA vs tensor DMMA (mma.sync m8n8k4 / m16n8k*) throughput on sm_100a,
// plus random 256-byte gather bandwidth (the access pattern of the BPMF Gram accumulate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 10 accumulators sharing 4 fragments (the upper-triangle-of-4x4-blocks pattern of the K=32 Gram)
__global__ void k_dmma884(double *out, int iters)
{
    double c[10][2];
#pragma unroll
    for (int i = 0; i < 10; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
        int t = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) { dmma884(c[t][0], c[t][1], f[a], f[b]); ++t; }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

#ifdef TRY_M16
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__global__ void k_dmma1684(double *out, int iters)
{
    double c[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    double f[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) f[i] = 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 6; ++t) dmma1684(c[t], f[t], f[(t + 1) % 6], f[(t + 2) % 6]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void k_dmma16816(double *out, int iters)
{
    double c[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 6; ++t) dmma16816(c[t], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
#endif

// random gather: every "rating" fetches one 256-byte latent vector (32 doubles).
// mode 0: a warp handles 4 ratings per step, lane l reads 32 B (4 doubles) of rating l%4 at offset (l/4)*32 B
// mode 1: a warp handles 1 rating per step, lane l reads 8 B
// mode 2: like mode 0 but unrolled x2 (8 ratings in flight per warp)
template <int MODE>
__global__ void k_gather(const double *__restrict__ tab, const int *__restrict__ idx, long n, double *out)
{
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    double s = 0;
    if (MODE == 1) {
        for (long p = warp * 32; p < n; p += nwarps * 32) {
            int my = idx[p + lane];
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                int j = __shfl_sync(0xffffffffu, my, r);
                s += tab[(long)j * 32 + lane];
            }
        }
    } else {
        for (long p = warp * 32; p < n; p += nwarps * 32) {
            int my = idx[p + lane];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                int j = __shfl_sync(0xffffffffu, my, g * 4 + (lane & 3));
                const double4 *src = reinterpret_cast<const double4 *>(tab + (long)j * 32 + (lane >> 2) * 4);
                double4 v = *src;
                s += v.x + v.y + v.z + v.w;
            }
        }
    }
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, nsm, prop.clockRate);
    double *out; CK(cudaMalloc(&out, 1024));
    const int iters = 20000;
    for (int wps : {4, 8, 16, 32}) {
        int threads = 128, blocks = nsm * wps / 4;
        float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 16 * iters * (double)blocks * threads;
        printf("DFMA       warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, iters / 4); });
        fl = 2.0 * 256 * 10 * (iters / 4) * (double)blocks * threads / 32;
        printf("DMMA m8n8k4 warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
#ifdef TRY_M16
        ms = time_ms([&] { k_dmma1684<<<blocks, threads>>>(out, iters / 4); });
        fl = 2.0 * 512 * 6 * (iters / 4) * (double)blocks * threads / 32;
        printf("DMMA m16n8k4 warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, iters / 16); });
        fl = 2.0 * 2048 * 6 * (iters / 16) * (double)blocks * threads / 32;
        printf("DMMA m16n8k16 warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
#endif
    }
    // gather: table of N vectors x 32 doubles
    for (long nvec : {1L << 18 /*64MB: L2 resident*/, 1L << 20 /*256MB*/, 1L << 22 /*1GB*/}) {
        double *tab; int *idx; long n = 1L << 26;  // 64M ratings -> 16 GiB gathered
        CK(cudaMalloc(&tab, nvec * 32 * sizeof(double))); CK(cudaMemset(tab, 0, nvec * 32 * sizeof(double)));
        std::vector<int> h(n);
        uint64_t s = 88172645463325252ull;
        for (long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)nvec); }
        CK(cudaMalloc(&idx, n * sizeof(int))); CK(cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
        for (int wps : {8, 16, 32, 64}) {
            int threads = 256, blocks = nsm * wps / 8;
            float ms0 = time_ms([&] { k_gather<0><<<blocks, threads>>>(tab, idx, n, out); }, 3);
            float ms1 = time_ms([&] { k_gather<1><<<blocks, threads>>>(tab, idx, n, out); }, 3);
            printf("gather table=%5ld MB warps/SM=%2d: mode0(4x8 lanes x32B) %7.1f GB/s   mode1(32 lanes x8B) %7.1f GB/s\n",
                   nvec * 256 >> 20, wps, n * 256.0 / ms0 * 1e-6, n * 256.0 / ms1 * 1e-6);
        }
        CK(cudaFree(tab)); CK(cudaFree(idx));
    }
    // streaming read for reference
    {
        long n = 1L << 29; double *buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
        double *dst; CK(cudaMalloc(&dst, n * 8));
        float ms = time_ms([&] { CK(cudaMemcpyAsync(dst, buf, n * 8, cudaMemcpyDeviceToDevice)); });
        printf("memcpy D2D 4 GiB: %.3f ms, %.1f GB/s (read+write)\n", ms, 2.0 * n * 8 / ms * 1e-6);
    }
    return 0;
}
