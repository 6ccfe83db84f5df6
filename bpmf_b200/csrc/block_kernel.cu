// block_kernel.cu — the fused per-item conditional update (c++/sample.cpp:263-336 + 248-258) for num_latent = 16 * m
// other than 32 (16, 48, 64, 80, 96, 112, 128) on sm_100a: ONE CTA PER ITEM, NB/2 warps (NB = K / 8).
//
//   gather  The K-vectors of the item's ratings are staged global -> shared with cp.async in stages of 32 rows (row pitch
//           8K + 32 bytes: conflict-free fragment loads), a ring of three stages (two in flight while one is consumed, ONE CTA
//           barrier per stage), every thread copying eight 16-byte chunks
//           whose row indices were loaded one stage earlier.
//   Gram    fp64 tensor cores (mma.sync.m8n8k4, DMMA). The lower triangle of NB x NB blocks of 8 x 8 is spread over the
//           warps' REGISTERS: warp w owns block rows w and NB-1-w (NB + 1 blocks, 2 NB + 2 accumulator doubles per
//           lane). Per four ratings it loads NB - w fragments and issues NB + 1 DMMAs; the rhs costs 2 DFMAs. Only that
//           inner part is instantiated per warp role (the operands of a DMMA must be registers); the stage loop around it
//           is shared code — with one copy of the whole loop per warp the kernel thrashed the instruction cache.
//   tail    MM = LambdaF + alpha G goes to shared memory as 8 x 8 row-major TILES of the lower block triangle (they alias
//           the gather ring, idle by then), with the right-hand side as one more block row. Blocked right-looking LDL^T
//           with look-ahead: warp 0 updates and factorises the next diagonal tile in registers (warp shuffles; the inverse
//           of its unit-lower factor comes out of the same row operations) while the other warps finish the trailing
//           update; a panel tile is then ONE 8x8x8 product with that inverse and the trailing tiles are DMMA updates
//           (fragments straight from the tiles: a row-major tile IS the C fragment layout, and its rows are A/B
//           fragments). Factorising the extra block row leaves D^-1 Lu^-1 b in it, i.e. the forward solve; the backward
//           solve (warp 0, vector in registers) is a shuffle + FMA chain; K normals from Philox4x32-10 (rng.cuh).
// Algorithmic bytes: 8K per rating. The Gram needs (NB+1) NB / 2 DMMAs per four ratings: at K = 128 that is 16.4 kflop per
// gathered KB, so the kernel is bound by the fp64 tensor pipe (37 TFLOP/s), not by HBM.
#include "common.cuh"
#include "rng.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace bpmf {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SR = 32;   // ratings per stage
constexpr int CPT = SR / 4;   // 16-byte chunks a thread copies per stage (the CTA has 2K threads, a stage 8K * SR / 16 chunks)
constexpr int NS = 3;    // stages in flight
#ifdef BPMF_BLOCK_PROF       // BPMF_BLOCK_PROF=1 python -m bpmf_b200.build --force: phase timestamps (BPMF_BLOCK_DBG=8)
constexpr bool BLOCK_PROF = true;
#else
constexpr bool BLOCK_PROF = false;
#endif

struct BlockArgs {
    int from, to;
    uint32_t iter;
    double alpha, mean_rating;
    const int64_t *colptr;
    const int32_t *rowidx;
    const double *val;
    const double *other;
    double *items;
    int npeers;
    double *const *peers;
    const double *mu, *LambdaF;
    unsigned int *work_counter;
    unsigned long long *err;
    const double *propLambda;   // per-item prior precisions K*K x num (-m / -l, sample.cpp:272-277) or nullptr
    // heavy items (more than heavy_thr ratings): their Gram was computed in chunks by block_gram_chunk_kernel
    int heavy_thr, n_heavy;
    const int *hv_item, *hv_first;      // item index (ascending), first chunk (n_heavy + 1)
    const double *hv_partials;          // per chunk: the accumulators of every thread, Cfg::PARTIAL doubles
    int dbg;   // timing probes (BPMF_BLOCK_DBG): 1 = no factorisation / solves, 2 = no Gram DMMAs, 4 = no solves, 8 = block 0 prints
               // the cycles of its phases for a few items; 0 = the product
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// The trailing-update schedules of every supported NB (build_schedules below), in constant memory: every lane of a warp
// reads the same entry.
constexpr int MAX_QUADS = 900, MAX_SPANS = 7 * 16 * 8;
__constant__ int4 c_ttab[2 * MAX_QUADS];
__constant__ int2 c_tspan[MAX_SPANS];
__constant__ int c_spanbase[17];          // [NB]: first entry of c_tspan for that NB

template <int NB>
struct Cfg {
    static constexpr int K = 8 * NB, NWB = NB / 2, T = 32 * NWB;
    static constexpr int ROWB = 8 * K + 32;                       // bytes per staged row
    static constexpr int STAGE = SR * ROWB;                       // rows only; weights live apart
    static constexpr int NTILE = NB * (NB + 1) / 2;               // 8 x 8 tiles of the lower block triangle, 512 B each
    // + block row NB (the right-hand side as one more row of the matrix: tile (NB, J) holds b[8J .. 8J+7] in its row 0)
    // + the inverses of the unit-lower diagonal tiles
    static constexpr int LINV = (NTILE + NB) * 64;                // doubles: inverse of diagonal tile kb at LINV + 64 kb
    static constexpr int NTILE_ALL = NTILE + 2 * NB;
    static constexpr int BIG = (NS * STAGE > NTILE_ALL * 512) ? NS * STAGE : NTILE_ALL * 512;   // ring and tiles alias each other
    // layout: [BIG][w: NS*SR][z: K][b: K][rr0: K][d: K][rinv: K][ints: 4]
    static constexpr int W_OFF = BIG, Z_OFF = W_OFF + NS * SR * 8, B_OFF = Z_OFF + K * 8, RR0_OFF = B_OFF + K * 8;
    static constexpr int D_OFF = RR0_OFF + K * 8, RI_OFF = D_OFF + K * 8, INT_OFF = RI_OFF + K * 8, SMEM = INT_OFF + 16;
    __host__ __device__ static constexpr int tile(int I, int J) { return (I * (I + 1) / 2 + J) * 64; }   // doubles
    static constexpr int PARTIAL = (2 * NB + 4) * T;             // doubles per chunk of a heavy item: acc[NB+1][2], r0, r1 of every thread
};

// Tile storage: element (r, c) of an 8 x 8 tile sits at 8 r + (c ^ swz(r)). Without the swizzle the fragment loads of a
// DMMA (lane (g, t) reads element (g, t) and (g, t + 4)) touch only half of the banks: 4 wavefronts instead of 2, and the
// tail of this kernel is bound by shared-memory wavefronts. Pairs (c, c + 1) with c even stay adjacent and 16-byte aligned.
__host__ __device__ __forceinline__ constexpr int swz(int r) { return (r & 2) << 1; }

// The Gram of the (at most SR) ratings of one stage for warp W (compile-time: the operands of a DMMA are registers):
// acc[0 .. W] = blocks (W, 0..W), acc[W+1 .. NB] = blocks (NB-1-W, 0..NB-1-W). Only this part of the per-item code differs
// between the warps; everything else is shared, which keeps the instruction footprint small.
template <int NB, int W>
__device__ __forceinline__ void gram_groups(double (&acc)[NB + 1][2], double &r0, double &r1, const unsigned char *stg, const double *swp,
                                            const int left, const int g, const int t, const int dbg)
{
    using C = Cfg<NB>;
    constexpr int R0 = W, R1 = NB - 1 - W, NF = NB - W;   // fragments f[0 .. NF-1] cover both rows
#pragma unroll 1
    for (int q = 0; q * 4 < left; ++q) {
        const unsigned char *row = stg + (4 * q + t) * C::ROWB + g * 8;
        double f[NF];
#pragma unroll
        for (int j = 0; j < NF; ++j) f[j] = *reinterpret_cast<const double *>(row + j * 64);
        const double w = swp[4 * q + t];
        if (!(dbg & 2)) {
#pragma unroll
            for (int j = 0; j <= R0; ++j) dmma884(acc[j][0], acc[j][1], f[R0], f[j]);
#pragma unroll
            for (int j = 0; j <= R1; ++j) dmma884(acc[R0 + 1 + j][0], acc[R0 + 1 + j][1], f[R1], f[j]);
        } else {
            acc[0][0] += f[0] * f[NF - 1];
        }
        r0 = fma(f[R0], w, r0);
        r1 = fma(f[R1], w, r1);
    }
}

// gather + Gram of the ratings [ps, pe) of one item (computeMuLambda, sample.cpp:251-257) by the whole CTA; adds to acc / r0 / r1
template <int NB>
__device__ __forceinline__ void gram_phase(const BlockArgs &p, unsigned char *smem, const int64_t ps, const int64_t pe, double (&acc)[NB + 1][2],
                                           double &r0, double &r1, const int tid, const int warp, const int g, const int t)
{
    using C = Cfg<NB>;
    constexpr int K = C::K, NWB = C::NWB;
            // ---- gather + Gram of the item (computeMuLambda, sample.cpp:251-257)
            const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(smem);
            double *sw = reinterpret_cast<double *>(smem + C::W_OFF);
            const int nst = (int)((pe - ps + SR - 1) / SR);
            // this thread's 16-byte chunks of a stage: rows rq, rq + 4, rq + 8, ..., chunk cq of the row
            const int rq = tid / (K / 2), cq = tid % (K / 2);
            unsigned pre[CPT];            // row indices of the NEXT stage to be issued, loaded one stage-iteration ahead
            auto load_rows = [&](int st) {
                const int64_t p0 = ps + (int64_t)st * SR;
#pragma unroll
                for (int j = 0; j < CPT; ++j) pre[j] = (p0 + rq + 4 * j < pe) ? (unsigned)__ldg(p.rowidx + p0 + rq + 4 * j) : 0u;
            };
            auto issue = [&](int st) {
                if (st < nst) {
                    const int64_t p0 = ps + (int64_t)st * SR;
                    const int slot = st % NS;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int r = rq + 4 * j;
                        cp_async16(ring_s + slot * C::STAGE + r * C::ROWB + cq * 16,
                                   reinterpret_cast<const unsigned char *>(p.other) + (size_t)pre[j] * (K * 8) + cq * 16, (p0 + r < pe) ? 16 : 0);
                    }
                    if (tid < SR) sw[slot * SR + tid] = (p0 + tid < pe) ? (__ldg(p.val + p0 + tid) - p.mean_rating) * p.alpha : 0.0;
                }
                cp_async_commit();
            };
#pragma unroll 1
            for (int st = 0; st < NS - 1; ++st) { load_rows(st); issue(st); }
            load_rows(NS - 1);
#pragma unroll 1
            for (int st = 0; st < nst; ++st) {
                // ONE barrier per stage: stage st has landed (for everyone), and everyone is done with stage st - 1, whose
                // slot is refilled right away with stage st + NS - 1 (so NS - 1 stages are in flight while st is consumed;
                // refilling after this stage's DMMAs instead measured the same)
                cp_async_wait<NS - 2>();
                __syncthreads();
                issue(st + NS - 1);
                load_rows(st + NS);
                const unsigned char *stg = smem + (st % NS) * C::STAGE;
                const double *swp = sw + (st % NS) * SR;
                const int left = (int)min((int64_t)SR, pe - (ps + (int64_t)st * SR));
                switch (warp) {           // the operands of a DMMA are registers: one small instantiation per warp role
                case 0: gram_groups<NB, 0>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 1: if constexpr (NWB > 1) gram_groups<NB, 1>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 2: if constexpr (NWB > 2) gram_groups<NB, 2>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 3: if constexpr (NWB > 3) gram_groups<NB, 3>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 4: if constexpr (NWB > 4) gram_groups<NB, 4>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 5: if constexpr (NWB > 5) gram_groups<NB, 5>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                case 6: if constexpr (NWB > 6) gram_groups<NB, 6>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                default: if constexpr (NWB > 7) gram_groups<NB, 7>(acc, r0, r1, stg, swp, left, g, t, p.dbg); break;
                }
            }
            cp_async_wait<0>();
            __syncthreads();              // the ring is dead
}

template <int NB>
__global__ void __launch_bounds__(Cfg<NB>::T, (NB >= 12 ? 2 : NB == 10 ? 3 : NB >= 6 ? 4 : 8)) items_block_kernel(BlockArgs p)
{
    using C = Cfg<NB>;
    constexpr int K = C::K, T = C::T, NWB = C::NWB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *MMp = reinterpret_cast<double *>(smem);
    double *z = reinterpret_cast<double *>(smem + C::Z_OFF), *b = reinterpret_cast<double *>(smem + C::B_OFF);
    double *rr0 = reinterpret_cast<double *>(smem + C::RR0_OFF);
    double *sd = reinterpret_cast<double *>(smem + C::D_OFF), *srinv = reinterpret_cast<double *>(smem + C::RI_OFF);
    const int g = lane >> 2, t = lane & 3;
    const int fpos = 8 * g + (t ^ swz(g));          // A / B fragment element (g, t); (g, t + 4) is at fpos ^ 4
    const int cpos = 8 * g + ((2 * t) ^ swz(g));    // C fragment pair (g, 2t), (g, 2t + 1)
    int *sint = reinterpret_cast<int *>(smem + C::INT_OFF);   // [0] item, [1] failed

    for (int a = tid; a < K; a += T) {    // LambdaF * hp.mu, the same for every item
        double s = 0.0;
        for (int j = 0; j < K; ++j) s += __ldg(p.LambdaF + a + (size_t)j * K) * __ldg(p.mu + j);
        rr0[a] = s;
    }
#pragma unroll 1
    for (;;) {
        __syncthreads();
        if (tid == 0) { sint[0] = p.from + (int)atomicAdd(p.work_counter, 1u); sint[1] = 0; }
        __syncthreads();
        const int idx = sint[0];
        if (idx >= p.to) break;
        const bool prof = BLOCK_PROF && (p.dbg & 8) && blockIdx.x == 0 && tid == 0;
        long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (prof) tk[0] = clock64();
        const double *LFi = p.propLambda ? p.propLambda + (size_t)idx * K * K : p.LambdaF;   // per-item prior precision (propagated posterior)
        if (p.propLambda) {               // rr = hp_LambdaF * hp.mu with the GLOBAL hp.mu (sample.cpp:285, quirk Q5)
            for (int a = tid; a < K; a += T) {
                double s2 = 0.0;
                for (int j = 0; j < K; ++j) s2 += __ldg(LFi + a + (size_t)j * K) * __ldg(p.mu + j);
                b[a] = s2;
            }
        } else {
            for (int a = tid; a < K; a += T) b[a] = rr0[a];
        }
        // the K normals of this item: rng_set_pos((idx+1) * K * (iter+1)) (sample.cpp:266)
        if (warp == NWB - 1) warp_randn((uint32_t)(((long long)idx + 1) * (long long)K * ((long long)p.iter + 1)), K, z);
        __syncthreads();
        const int64_t ps = __ldg(p.colptr + idx), pe = __ldg(p.colptr + idx + 1);
        {
            double acc[NB + 1][2], r0 = 0.0, r1 = 0.0;
#pragma unroll
            for (int n = 0; n <= NB; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
            if (pe - ps > p.heavy_thr && p.n_heavy > 0) {
                // a heavy item: its ratings were cut into chunks whose partial Grams block_gram_chunk_kernel computed; they
                // are added in chunk order, so the result does not depend on scheduling
                int lo = 0, hi = p.n_heavy - 1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(p.hv_item + mid) < idx) lo = mid + 1; else hi = mid; }
                for (int ch = __ldg(p.hv_first + lo); ch < __ldg(p.hv_first + lo + 1); ++ch) {
                    const double *in = p.hv_partials + (size_t)ch * C::PARTIAL + tid;
#pragma unroll
                    for (int n = 0; n <= NB; ++n) { acc[n][0] += in[(2 * n) * T]; acc[n][1] += in[(2 * n + 1) * T]; }
                    r0 += in[(2 * NB + 2) * T]; r1 += in[(2 * NB + 3) * T];
                }
            } else {
                gram_phase<NB>(p, smem, ps, pe, acc, r0, r1, tid, warp, g, t);
            }
            __syncthreads();              // the ring is dead: the tiles may overwrite it
            // MM = LambdaF + alpha * G (sample.cpp:297-298) into the tiles: lane (g, t) owns elements (g, 2t), (g, 2t + 1).
            // acc[n] is block (warp, n) for n <= warp and block (NB-1-warp, n-warp-1) after that.
#pragma unroll
            for (int n = 0; n <= NB; ++n) {
                const int I = (n <= warp) ? warp : NB - 1 - warp, J = (n <= warp) ? n : n - warp - 1;
                const int i = 8 * I + g, k = 8 * J + 2 * t;
                double2 v;
                v.x = fma(p.alpha, acc[n][0], __ldg(LFi + i + (size_t)k * K));
                v.y = fma(p.alpha, acc[n][1], __ldg(LFi + i + (size_t)(k + 1) * K));
                *reinterpret_cast<double2 *>(MMp + C::tile(I, J) + cpos) = v;
            }
            // rhs: sum over the quad's four ratings-of-a-group, on top of LambdaF * mu (sample.cpp:285) already in b
            r0 += __shfl_xor_sync(FULL, r0, 1); r0 += __shfl_xor_sync(FULL, r0, 2);
            r1 += __shfl_xor_sync(FULL, r1, 1); r1 += __shfl_xor_sync(FULL, r1, 2);
            if (t == 0) { b[8 * warp + g] += r0; b[8 * (NB - 1 - warp) + g] += r1; }
        }
        __syncthreads();
        if (prof) tk[1] = clock64();
        if (p.dbg & 1) {
            if (warp == 0 && lane < 8) p.items[(size_t)idx * K + lane] = MMp[lane] + b[lane] + z[lane];
            continue;
        }
        // ---- chol.compute(MM) (sample.cpp:306) as MM = Lu D Lu^T, blocked by 8, on the tiles, with look-ahead: the
        // diagonal tile of block column kb + 1 is updated and factorised by warp 0 while the other warps run the rest of
        // the trailing update of block column kb, so its 8-pivot dependent chain is off the other warps' critical path.
        // factor_diag: EVERY lane of the warp holds the whole lower triangle of the tile in registers and runs the same
        // fully unrolled 8 x 8 factorisation — no shuffles, the pivot chain is rcp seed -> 2 DFMA -> multiplier -> update,
        // everything else is independent work. The inverse of the unit-lower factor (the panel's operand) comes out of
        // the same row operations applied to I. All lanes store the same values (the tile is lane-uniform).
        // The right-hand side is block row NB of the matrix: row 0 of tile (NB, J) = b[8J .. 8J+7], rows 1..7 zero.
        // Factorising the matrix with this extra row leaves D^-1 Lu^-1 b in it (the forward solve, sample.cpp:321).
        for (int e = tid; e < NB * 64; e += T) {
            MMp[C::tile(NB, 0) + e] = ((e & 63) < 8) ? b[(e >> 6) * 8 + (e & 7)] : 0.0;
            const int er = (e >> 3) & 7, ec = e & 7;                                 // diagonal and upper part of the inverses
            if (ec >= er) MMp[C::LINV + (e & ~7) + (ec ^ swz(er))] = (ec == er) ? 1.0 : 0.0;
        }
        auto factor_diag = [&](const int kb) {
            double *tp = MMp + C::tile(kb, kb), *lp = MMp + C::LINV + 64 * kb;
            double a[8][8], W[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j <= i; j += 2) {
                    const double2 v2 = *reinterpret_cast<const double2 *>(tp + 8 * i + (j ^ swz(i)));
                    a[i][j] = v2.x;
                    if (j + 1 <= i) a[i][j + 1] = v2.y;
                }
            }
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = a[k][k];
                ok = ok && (d > 0.0);                                                // pivot <= 0 -> "Cholesky failed"
                double r;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
                const double e1 = fma(-d, r, 1.0);
                const double e2 = fma(e1, e1, e1);                                   // 1 / d = r + r e2
                if (lane == 0) { sd[8 * kb + k] = d; srinv[8 * kb + k] = fma(r, e2, r); }
                double l[8];
#pragma unroll
                for (int i = k + 1; i < 8; ++i) { const double u = a[i][k] * r; l[i] = fma(u, e2, u); }   // Lu(i,k) = a(i,k) / d
#pragma unroll
                for (int j = k + 1; j < 8; ++j) {
#pragma unroll
                    for (int i = j; i < 8; ++i) a[i][j] = fma(-l[i], a[j][k], a[i][j]);
                }
#pragma unroll
                for (int i = k + 1; i < 8; ++i) {                                    // W <- (I - l e_k^T) W
#pragma unroll
                    for (int j = 0; j < k; ++j) W[i][j] = fma(-l[i], W[k][j], W[i][j]);
                    W[i][k] = -l[i];
                    tp[8 * i + (k ^ swz(i))] = l[i];
                }
            }
#pragma unroll
            for (int i = 1; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < i; ++j) lp[8 * i + (j ^ swz(i))] = W[i][j];
            }
            if (!ok && lane == 0) sint[1] = 1;
        };
        if (warp == 0) factor_diag(0);
        __syncthreads();
        if (prof) tk[2] = clock64();
        const bool prof1 = BLOCK_PROF && (p.dbg & 8) && blockIdx.x == 0 && tid == 32;
        long long w1_trail = 0, w1_first = 0;
        int w1_tiles = 0;
#pragma unroll 1
        for (int kb = 0; kb < NB; ++kb) {
            long long c0 = 0, c1 = 0, c2 = 0;
            if (prof) c0 = clock64();
            // ---- panel: this warp's tiles below the (already factorised) diagonal tile, I = I0, I0 + NWB (block row NB is
            // the right-hand side):  Lu(I,kb) = A(I,kb) Lu(kb,kb)^-T D^-1, one 8x8x8 product on the tensor cores
            const int I0 = kb + 1 + ((warp - (kb + 1)) % NWB + NWB) % NWB;
            if (I0 <= NB) {
                const double *li = MMp + C::LINV + 64 * kb;                 // B fragment: B[k][n] = Linv[n][k]
                const double lb0 = li[fpos], lb1 = li[fpos ^ 4];
                const double rv0 = srinv[8 * kb + 2 * t], rv1 = srinv[8 * kb + 2 * t + 1];
                const bool two = I0 + NWB <= NB;
                double *tp0 = MMp + C::tile(I0, kb), *tp1 = MMp + C::tile(two ? I0 + NWB : I0, kb);
                const double a00 = tp0[fpos], a01 = tp0[fpos ^ 4], a10 = tp1[fpos], a11 = tp1[fpos ^ 4];
                double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;
                dmma884(c00, c01, a00, lb0);   // four independent DMMAs: the k-halves are added afterwards
                dmma884(c10, c11, a10, lb0);
                dmma884(e00, e01, a01, lb1);
                dmma884(e10, e11, a11, lb1);
                __syncwarp();                  // every lane has read the tiles before they are overwritten
                *reinterpret_cast<double2 *>(tp0 + cpos) = make_double2((c00 + e00) * rv0, (c01 + e01) * rv1);
                if (two) *reinterpret_cast<double2 *>(tp1 + cpos) = make_double2((c10 + e10) * rv0, (c11 + e11) * rv1);
            }
            __syncthreads();
            if (prof) c1 = clock64();
            // ---- trailing tiles (I, J), kb < J <= I:  A(I,J) -= Lu(I,kb) D Lu(J,kb)^T  on the tensor cores
            if (kb + 1 < NB) {
                const double dc0 = sd[8 * kb + t], dc1 = sd[8 * kb + 4 + t];
                if (warp == 0) {               // the next diagonal tile first, then its factorisation
                    double *tc = MMp + C::tile(kb + 1, kb + 1) + cpos;
                    const double *ta = MMp + C::tile(kb + 1, kb);
                    const double ta0 = ta[fpos], ta1 = ta[fpos ^ 4];
                    double2 cv = *reinterpret_cast<const double2 *>(tc);
                    double y0 = 0.0, y1 = 0.0;
                    dmma884(cv.x, cv.y, -ta0, ta0 * dc0);
                    dmma884(y0, y1, -ta1, ta1 * dc1);
                    *reinterpret_cast<double2 *>(tc) = make_double2(cv.x + y0, cv.y + y1);
                    __syncwarp();
                    factor_diag(kb + 1);
                    if (prof) c2 = clock64();
                }
                if (warp > 0 || NWB == 1) {
                    // the other tiles (rest of the trailing triangle, then block row NB = the right-hand side), dealt
                    // round-robin to the warps by a schedule built on the host (it does not depend on the item). U tiles
                    // in flight per warp; the two k-halves of a tile go to separate accumulators so that no DMMA waits
                    // for another one.
                    // schedule: two int4 per QUAD (offsets of Lu(I0,kb), Lu(I1,kb), Lu(J0,kb), Lu(J1,kb); of the four tiles, -1 = none)
                    const int2 span = c_tspan[c_spanbase[NB] + kb * NWB + warp];
                    long long q0 = 0;
                    if (prof1) q0 = clock64();
                    int4 ea = make_int4(0, 0, 0, 0), ec = make_int4(-1, -1, -1, -1);
                    if (span.x < span.y) { ea = c_ttab[2 * span.x]; ec = c_ttab[2 * span.x + 1]; }
#pragma unroll 1
                    for (int n = span.x; n < span.y; ++n) {
                        // a QUAD of tiles (I0|I1, J0|J1): two A and two B fragments pairs serve four tiles
                        const int4 qa = ea, qc = ec;
                        if (n + 1 < span.y) { ea = c_ttab[2 * n + 2]; ec = c_ttab[2 * n + 3]; }
                        const double a00 = MMp[qa.x + fpos], a01 = MMp[qa.x + (fpos ^ 4)], a10 = MMp[qa.y + fpos], a11 = MMp[qa.y + (fpos ^ 4)];
                        const double b00 = MMp[qa.z + fpos] * dc0, b01 = MMp[qa.z + (fpos ^ 4)] * dc1;
                        const double b10 = MMp[qa.w + fpos] * dc0, b11 = MMp[qa.w + (fpos ^ 4)] * dc1;
                        const int co[4] = {qc.x, qc.y, qc.z, qc.w};               // (I0,J0) (I0,J1) (I1,J0) (I1,J1); < 0: not a tile
                        double2 cv[4];
                        double x0[4], x1[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            cv[u] = co[u] >= 0 ? *reinterpret_cast<const double2 *>(MMp + co[u] + cpos) : make_double2(0.0, 0.0);
                            x0[u] = 0.0; x1[u] = 0.0;
                        }
                        dmma884(cv[0].x, cv[0].y, -a00, b00);
                        dmma884(cv[1].x, cv[1].y, -a00, b10);
                        dmma884(cv[2].x, cv[2].y, -a10, b00);
                        dmma884(cv[3].x, cv[3].y, -a10, b10);
                        dmma884(x0[0], x1[0], -a01, b01);
                        dmma884(x0[1], x1[1], -a01, b11);
                        dmma884(x0[2], x1[2], -a11, b01);
                        dmma884(x0[3], x1[3], -a11, b11);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (co[u] >= 0) {
                                *reinterpret_cast<double2 *>(MMp + co[u] + cpos) = make_double2(cv[u].x + x0[u], cv[u].y + x1[u]);
                                if (prof1) ++w1_tiles;
                            }
                    }
                    if (prof1) { const long long q1 = clock64(); w1_trail += q1 - q0; if (kb == 0) w1_first = q1 - q0; }
                }
            }
            __syncthreads();
            if (prof) { const long long c3 = clock64(); tk[3] += c1 - c0; if (c2) { tk[4] += c2 - c1; tk[5] += c3 - c2; } else tk[5] += c3 - c1; }
        }
        if (prof) tk[6] = clock64();
        if (prof1 && idx < p.from + 1500) printf("  warp 1 of item %d: %d trailing tiles in %lld cycles (block column 0: %lld)\n", idx, w1_tiles, w1_trail, w1_first);
        if (sint[1]) {                    // THROWERROR("Cholesky failed") (sample.cpp:308)
            if (tid == 0) atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
            continue;
        }
        // ---- L \ rr; rr += nrandn; L^T \ rr (sample.cpp:321-323) with L = Lu D^(1/2):  x = Lu^-T (D^-1 Lu^-1 b + D^(-1/2) z).
        // D^-1 Lu^-1 b is row 0 of block row NB; the backward solve runs on warp 0 with the vector in registers: lane holds
        // rows lane + 32 r. Lu(i,k) = tile(i / 8, k / 8)[8 (i % 8) + k % 8].
        if (warp == 0 && !(p.dbg & 4)) {
            constexpr int R = (K + 31) / 32;
            double v[R];
            int colo[R];                  // this lane's rows i = lane + 32 r as COLUMN offsets inside a tile row
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane + 32 * r, I = i >> 3;
                colo[r] = I * 64 + (i & 7);
                v[r] = (i < K) ? fma(rsqrt(sd[i]), z[i], MMp[C::tile(NB, I) + (i & 7)]) : 0.0;
            }
            // one tile row (8 columns) at a time: its Lu values are loaded up front, so that the chain per column is
            // shuffle -> DFMA only
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
#pragma unroll 1
                for (int q = 3; q >= 0; --q) {
                    const int kb = 4 * r + q;
                    if (kb >= NB) continue;
                    const double *krow = MMp + (kb * (kb + 1) / 2) * 64;          // Lu(8 kb + c, i) = krow[8 c + colo(i)]
                    double Lv[8][R];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
#pragma unroll
                        for (int r2 = 0; r2 <= r; ++r2) Lv[c][r2] = krow[8 * c + (colo[r2] ^ swz(c))];
                    }
#pragma unroll
                    for (int c = 7; c >= 0; --c) {
                        const int k = 8 * kb + c;
                        const double xk = __shfl_sync(FULL, v[r], 8 * q + c);
#pragma unroll
                        for (int r2 = 0; r2 <= r; ++r2) {
                            if (lane + 32 * r2 < k) v[r2] = fma(-Lv[c][r2], xk, v[r2]);
                        }
                    }
                }
            }
            if (prof && idx < p.from + 1500)
                printf("item %d nnz %d: gram %lld  init+diag0 %lld  panels %lld  diag %lld  wait_trailing %lld  solves %lld\n", idx, (int)(pe - ps),
                       tk[1] - tk[0], tk[2] - tk[1], tk[3], tk[4], tk[5], clock64() - tk[6]);
            // items().col(idx) = rr (sample.cpp:324), and into every peer replica (replaces send_item)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane + 32 * r;
                if (i < K) {
                    p.items[(size_t)idx * K + i] = v[r];
                    for (int q = 0; q < p.npeers; ++q)
                        if (p.peers[q] && p.peers[q] != p.items) p.peers[q][(size_t)idx * K + i] = v[r];
                }
            }
        }
    }
}

// One CTA per chunk of a heavy item: the partial Gram and right-hand side of its ratings, in the accumulator layout of
// items_block_kernel (every thread's acc[NB+1][2], r0, r1), to be added there in chunk order.
template <int NB>
__global__ void __launch_bounds__(Cfg<NB>::T, (NB >= 12 ? 2 : NB == 10 ? 3 : NB >= 6 ? 4 : 8))
    block_gram_chunk_kernel(BlockArgs p, const int64_t *__restrict__ ch_p0, const int64_t *__restrict__ ch_p1, double *__restrict__ partials)
{
    using C = Cfg<NB>;
    constexpr int T = C::T;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ch = blockIdx.x;
    double acc[NB + 1][2], r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int n = 0; n <= NB; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
    gram_phase<NB>(p, smem, ch_p0[ch], ch_p1[ch], acc, r0, r1, tid, warp, g, t);
    double *out = partials + (size_t)ch * C::PARTIAL + tid;
#pragma unroll
    for (int n = 0; n <= NB; ++n) { out[(2 * n) * T] = acc[n][0]; out[(2 * n + 1) * T] = acc[n][1]; }
    out[(2 * NB + 2) * T] = r0; out[(2 * NB + 3) * T] = r1;
}

// The trailing-update schedule of block column kb: the trailing triangle plus block row NB (the right-hand side), minus
// the next diagonal tile (warp 0's), cut into 2 x 2 QUADS of tiles that share their A / B fragments; the quads go
// round-robin to warps 1 .. NWB-1 (to the only warp when NWB == 1). It does not depend on the item.
struct Schedules {
    std::vector<int4> tab;
    std::vector<int2> span;
    int spanbase[17] = {0};
};

static Schedules make_schedules()
{
    Schedules S;
    auto tile = [](int I, int J) { return (I * (I + 1) / 2 + J) * 64; };
    for (int NB : {2, 6, 8, 10, 12, 14, 16}) {
        const int NWB = NB / 2;
        S.spanbase[NB] = (int)S.span.size();
        std::vector<std::vector<int4>> per((size_t)NB * NWB);        // two int4 per quad
        for (int kb = 0; kb + 1 < NB; ++kb) {
            // a tile (I, J) of this block column's trailing update: kb < J <= I < NB, or block row NB (the right-hand
            // side) with kb < J < NB; the next diagonal tile is warp 0's
            auto live = [&](int I, int J) {
                if (J <= kb || J >= NB || I > NB || J > I) return false;
                return !(I == kb + 1 && J == kb + 1);
            };
            int q = 0;
            for (int I0 = kb + 1; I0 <= NB; I0 += 2)
                for (int J0 = kb + 1; J0 < NB && J0 <= I0 + 1; J0 += 2) {
                    const int I1 = I0 + 1, J1 = J0 + 1;
                    const bool l00 = live(I0, J0), l01 = live(I0, J1), l10 = live(I1, J0), l11 = live(I1, J1);
                    if (!(l00 || l01 || l10 || l11)) continue;
                    const int w = NWB > 1 ? 1 + q % (NWB - 1) : 0;
                    ++q;
                    const int a0 = tile(I0, kb), a1 = I1 <= NB ? tile(I1, kb) : a0;
                    const int b0 = tile(J0, kb), b1 = J1 < NB ? tile(J1, kb) : b0;
                    auto &v = per[(size_t)kb * NWB + w];
                    v.push_back(make_int4(a0, a1, b0, b1));
                    v.push_back(make_int4(l00 ? tile(I0, J0) : -1, l01 ? tile(I0, J1) : -1, l10 ? tile(I1, J0) : -1, l11 ? tile(I1, J1) : -1));
                }
        }
        for (size_t i = 0; i < per.size(); ++i) {
            S.span.push_back(make_int2((int)(S.tab.size() / 2), (int)((S.tab.size() + per[i].size()) / 2)));    // in quads
            S.tab.insert(S.tab.end(), per[i].begin(), per[i].end());
        }
    }
    return S;
}

static cudaError_t build_schedules(int device)
{
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lock(mu);
    for (int d : done) if (d == device) return cudaSuccess;
    const Schedules S = make_schedules();
    if (S.tab.size() > 2 * (size_t)MAX_QUADS || S.span.size() > (size_t)MAX_SPANS) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemcpyToSymbol(c_ttab, S.tab.data(), sizeof(int4) * S.tab.size());
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_tspan, S.span.data(), sizeof(int2) * S.span.size());
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_spanbase, S.spanbase, sizeof S.spanbase);
    if (e != cudaSuccess) return e;
    done.push_back(device);
    return cudaSuccess;
}

template <int NB>
cudaError_t launch_nb(bpmf_gpu_ctx *c, const BlockArgs &p, long long n, int ch0, int ch1, const int64_t *hv_p0, const int64_t *hv_p1, double *hv_partials)
{
    using C = Cfg<NB>;
    static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
    if (ch1 > ch0) {                      // the chunks of the heavy items of [from, to) first, on the same stream
        auto ck = block_gram_chunk_kernel<NB>;
        cudaError_t ec = cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (ec != cudaSuccess) return ec;
        ck<<<(unsigned)(ch1 - ch0), C::T, C::SMEM, c->stream>>>(p, hv_p0 + ch0, hv_p1 + ch0, hv_partials + (size_t)ch0 * C::PARTIAL);
        c->launches++;
        if ((ec = cudaGetLastError()) != cudaSuccess) return ec;
    }
    auto kern = items_block_kernel<NB>;
    // BPMF_BLOCK_1CTA (timing probe): pad the shared memory request so that only one CTA fits an SM
    static const bool one_cta = getenv("BPMF_BLOCK_1CTA") != nullptr;
    const int smem = one_cta && C::SMEM < 120 * 1024 ? 120 * 1024 : C::SMEM;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    if (grid > n) grid = n;
    kern<<<(unsigned)grid, C::T, smem, c->stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

// doubles of a heavy item's chunk partial for the CTA-per-item kernel (capi.cu sizes hv_partials with it)
int block_partial_doubles(int K)
{
    const int NB = K / 8;
    return (2 * NB + 4) * 32 * (NB / 2);
}
int block_heavy_chunk_size() { return 4096; }

bool block_kernel_supports(int K) { return K % 16 == 0 && K >= 16 && K <= 128 && K != 32; }

// Host-only view of the trailing-update schedule (tests): the quads warp `warp` takes in block column kb, eight ints per
// quad (tile offsets in doubles: Lu(I0,kb), Lu(I1,kb), Lu(J0,kb), Lu(J1,kb), A(I0,J0), A(I0,J1), A(I1,J0), A(I1,J1); -1 =
// not a tile). Returns the number of quads, or -1.
int block_schedule(int K, int kb, int warp, int *out, int cap_quads)
{
    if (!block_kernel_supports(K)) return -1;
    const int NB = K / 8, NWB = NB / 2;
    if (kb < 0 || kb >= NB || warp < 0 || warp >= NWB) return -1;
    static const Schedules S = make_schedules();
    const int2 sp = S.span[(size_t)S.spanbase[NB] + (size_t)kb * NWB + warp];
    const int n = sp.y - sp.x;
    for (int q = 0; q < n && q < cap_quads; ++q) {
        const int4 a = S.tab[2 * (size_t)(sp.x + q)], c = S.tab[2 * (size_t)(sp.x + q) + 1];
        const int v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        for (int i = 0; i < 8; ++i) out[8 * q + i] = v[i];
    }
    return n;
}

cudaError_t launch_items_block(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    BlockArgs p;
    p.from = s.from; p.to = s.to; p.iter = iter; p.alpha = alpha; p.mean_rating = s.mean_rating;
    p.colptr = s.colptr; p.rowidx = s.rowidx; p.val = s.val;
    p.other = o.items; p.items = s.items;
    p.npeers = s.npeers; p.peers = s.peers_dev;
    p.mu = s.hp.mu; p.LambdaF = s.hp.LambdaF;
    p.work_counter = s.work_counter; p.err = c->d_err;
    p.propLambda = s.propLambda;
    // heavy items of [from, to): their chunks are reduced by block_gram_chunk_kernel before the item kernel
    int first = 0, last = 0;
    while (first < s.n_heavy && s.h_heavy_item[(size_t)first] < s.from) ++first;
    last = first;
    while (last < s.n_heavy && s.h_heavy_item[(size_t)last] < s.to) ++last;
    p.heavy_thr = s.heavy_thr; p.n_heavy = s.n_heavy; p.hv_item = s.hv_item; p.hv_first = s.hv_first; p.hv_partials = s.hv_partials;
    const int ch0 = s.n_heavy ? s.h_heavy_first[(size_t)first] : 0, ch1 = s.n_heavy ? s.h_heavy_first[(size_t)last] : 0;
    static const int dbg = [] { const char *v = getenv("BPMF_BLOCK_DBG"); return v ? atoi(v) : 0; }();
    p.dbg = dbg;
    cudaError_t e = cudaSuccess;
    if ((e = build_schedules(c->device)) != cudaSuccess) return e;     // once per device
    e = cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned int), c->stream);
    if (e != cudaSuccess) return e;
    const long long n = (long long)s.to - s.from;
    if (n < 1) return cudaSuccess;
    switch (c->K) {
    case 16: e = launch_nb<2>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 48: e = launch_nb<6>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 64: e = launch_nb<8>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 80: e = launch_nb<10>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 96: e = launch_nb<12>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 112: e = launch_nb<14>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    case 128: e = launch_nb<16>(c, p, n, ch0, ch1, s.hv_p0, s.hv_p1, s.hv_partials); break;
    default: return cudaErrorInvalidValue;
    }
    c->launches++;
    return e;
}

}  // namespace bpmf
