// Microbenchmark: is the fp64 pipe's time ADDITIVE when DMMAs and scalar DFMAs of different warps (or of one warp)
// are interleaved on one scheduler of sm_100a, or does every switch between the two cost dead cycles?
// One CTA per SM; warp w runs on scheduler w % 4. Roles per warp:
//   D  streams nD independent DMMAs (10 accumulators, like the Gram of the BPMF kernel)
//   F  streams nF independent DFMAs (8 chains)
//   C  one dependent DFMA chain of nF
//   S  dependent chain of 64-bit SHFL + DFMA (a triangular-solve step), nF steps
//   M  one warp alternating bursts: 10 DMMAs then `mix` independent DFMAs (the Gram group pattern), nD DMMAs in total
// Reported: elapsed cycles of the whole CTA (clock64, max over warps) per scheduler, against the additive model
//   nD * 16 + nF * cF  with cF measured by the F-only run.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Cfg { char role[32]; int nwarps; int nD, nF, mix; };

__global__ void k_mix(Cfg cfg, long long *cycles, double *sink)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const char role = cfg.role[warp];
    double c[10][2];
#pragma unroll
    for (int i = 0; i < 10; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-9 * (lane + i);
    const double a = 1.0000001, b = 1e-9, f = 1e-3 * lane;
    __syncthreads();
    const long long t0 = clock64();
    if (role == 'D') {
        for (int i = 0; i < cfg.nD; i += 10) {
#pragma unroll
            for (int j = 0; j < 10; ++j) dmma884(c[j][0], c[j][1], f, f);
        }
    } else if (role == 'F') {
        for (int i = 0; i < cfg.nF; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        }
    } else if (role == 'C') {
#pragma unroll 8
        for (int i = 0; i < cfg.nF; ++i) x[0] = fma(x[0], a, b);
    } else if (role == 'S') {
#pragma unroll 8
        for (int i = 0; i < cfg.nF; ++i) { const double t = __shfl_sync(0xffffffffu, x[0], i & 31); x[0] = fma(-a, t, x[0]); }
    } else if (role == 'M') {
        // (x[] must be indexed by compile-time constants: a run-time index would put it in local memory)
#define BURSTS(MIX)                                                                  \
        for (int i = 0; i < cfg.nD; i += 10) {                                       \
            _Pragma("unroll") for (int j = 0; j < 10; ++j) dmma884(c[j][0], c[j][1], f, f); \
            _Pragma("unroll") for (int j = 0; j < MIX; ++j) x[j & 7] = fma(x[j & 7], a, b); \
        }
        if (cfg.mix == 0) { BURSTS(0) } else if (cfg.mix == 4) { BURSTS(4) } else if (cfg.mix == 8) { BURSTS(8) }
        else if (cfg.mix == 16) { BURSTS(16) } else { BURSTS(32) }
#undef BURSTS
    } else if (role == 'N') {
        // the same bursts with DEPENDENT DFMAs (one chain), i.e. a tail step between Gram groups of the same warp
#define BURSTS(MIX)                                                                  \
        for (int i = 0; i < cfg.nD; i += 10) {                                       \
            _Pragma("unroll") for (int j = 0; j < 10; ++j) dmma884(c[j][0], c[j][1], f, f); \
            _Pragma("unroll") for (int j = 0; j < MIX; ++j) x[0] = fma(x[0], a, b); \
        }
        if (cfg.mix == 4) { BURSTS(4) } else if (cfg.mix == 8) { BURSTS(8) } else { BURSTS(16) }
#undef BURSTS
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) sink[0] = s;
}

static double run(const char *roles, int nD, int nF, int mix, long long *d_cyc, double *d_sink, double *per_role_out = nullptr)
{
    Cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.nwarps = (int)strlen(roles);
    memcpy(cfg.role, roles, cfg.nwarps);
    cfg.nD = nD; cfg.nF = nF; cfg.mix = mix;
    static long long h[148 * 32];
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(d_cyc, 0, sizeof(h)));
        k_mix<<<148, cfg.nwarps * 32>>>(cfg, d_cyc, d_sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int w = 0; w < cfg.nwarps; ++w) mx = h[w] > mx ? (double)h[w] : mx;   // block 0
        if (mx < best) {
            best = mx;
            if (per_role_out)
                for (int w = 0; w < cfg.nwarps; ++w) per_role_out[w] = (double)h[w];
        }
    }
    return best;
}

int main()
{
    long long *cyc; double *sink;
    CK(cudaMalloc(&cyc, 148 * 32 * 8)); CK(cudaMalloc(&sink, 64));
    const int nD = 4000, nF = 8000;
    double pw[32];
    // one role per scheduler: warps 0..3 (and 4..7 ...) land on schedulers 0..3
    const double tD = run("DDDD", nD, nF, 0, cyc, sink);
    const double tF = run("FFFF", nD, nF, 0, cyc, sink);
    const double tC = run("CCCC", nD, nF, 0, cyc, sink);
    const double tS = run("SSSS", nD, nF, 0, cyc, sink);
    printf("alone, one warp per scheduler: DMMA %.2f cycles each | independent DFMA %.2f | dependent DFMA %.2f | SHFL+DFMA step %.2f\n",
           tD / nD, tF / nF, tC / nF, tS / nF);
    struct { const char *name, *roles; int nF; } cases[] = {
        {"1 D + 1 F per scheduler", "DDDDFFFF", nF},
        {"1 D + 2 F per scheduler", "DDDDFFFFFFFF", nF},
        {"2 D + 1 F per scheduler", "DDDDDDDDFFFF", nF},
        {"2 D + 2 F per scheduler", "DDDDDDDDFFFFFFFF", nF},
        {"1 D + 1 C per scheduler (dependent chain)", "DDDDCCCC", nF / 4},
        {"2 D + 1 C per scheduler", "DDDDDDDDCCCC", nF / 4},
        {"2 D + 3 C per scheduler", "DDDDDDDDCCCCCCCCCCCC", nF / 4},
        {"3 D + 2 C per scheduler", "DDDDDDDDDDDDCCCCCCCC", nF / 4},
        {"2 D + 3 S per scheduler (solve steps)", "DDDDDDDDSSSSSSSSSSSS", nF / 8},
        {"0 D + 5 C per scheduler", "CCCCCCCCCCCCCCCCCCCC", nF / 4},
        {"0 D + 5 S per scheduler", "SSSSSSSSSSSSSSSSSSSS", nF / 8},
    };
    for (auto &cs : cases) {
        const int nw = (int)strlen(cs.roles);
        const double t = run(cs.roles, nD, cs.nF, 0, cyc, sink, pw);
        int nDw = 0, nFw = 0;
        double tDmax = 0, tFmax = 0;
        for (int w = 0; w < nw; w += 4) {          // the warps of scheduler 0
            if (cs.roles[w] == 'D') { ++nDw; tDmax = pw[w] > tDmax ? pw[w] : tDmax; } else { ++nFw; tFmax = pw[w] > tFmax ? pw[w] : tFmax; }
        }
        const double cF = tF / nF;
        const double additive = nDw * nD * (tD / nD) + nFw * cs.nF * cF;
        printf("%-44s elapsed %9.0f cycles | D warps done at %9.0f, others at %9.0f | additive pipe-time model %9.0f -> %.2fx | per scalar instr of one warp %.1f cycles\n",
               cs.name, t, tDmax, tFmax, additive, t / additive, tFmax / cs.nF);
    }
    // one warp per scheduler alternating 10 DMMAs and `mix` DFMAs (the Gram group: 10 DMMA + 4 DFMA)
    for (int mix : {0, 4, 8, 16, 32}) {
        const double t = run("MMMM", nD, nF, mix, cyc, sink);
        const double add = nD * (tD / nD) + (nD / 10) * mix * (tF / nF);
        printf("one warp, 10 DMMA + %2d DFMA bursts: %9.0f cycles, additive model %9.0f -> %.2fx\n", mix, t, add, t / add);
    }
    for (int mix : {4, 8, 16}) {
        const double t = run("NNNN", nD, nF, mix, cyc, sink);
        const double add = nD * (tD / nD) + (nD / 10) * mix * (tF / nF);
        printf("one warp, 10 DMMA + %2d DEPENDENT DFMA bursts: %9.0f cycles, additive model %9.0f -> %.2fx (%.1f cycles per DFMA beyond the DMMAs)\n", mix, t,
               add, t / add, (t - nD * (tD / nD)) / ((nD / 10) * mix));
    }
    for (int mix : {4, 16}) {
        const double t = run("MMMMMMMM", nD, nF, mix, cyc, sink);
        const double add = 2 * (nD * (tD / nD) + (nD / 10) * mix * (tF / nF));
        printf("two warps per scheduler, 10 DMMA + %2d DFMA bursts: %9.0f cycles, additive model %9.0f -> %.2fx\n", mix, t, add, t / add);
    }
    return 0;
}
