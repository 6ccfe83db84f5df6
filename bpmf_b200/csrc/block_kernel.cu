// block_kernel.cu — the fused per-item conditional update (c++/sample.cpp:263-336 + 248-258) for num_latent = 16 * m
// other than 32 (16, 48, 64, 80, 96, 112, 128) on sm_100a: ONE CTA PER ITEM, NB/2 warps (NB = K / 8).
//
//   gather  The K-vectors of the item's ratings are staged global -> shared with cp.async in stages of 32 rows (row pitch
//           8K + 32 bytes: conflict-free fragment loads), three stages in flight, every thread copying eight 16-byte chunks
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

#include <cstdlib>

namespace bpmf {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SR = 32;   // ratings per stage
constexpr int CPT = SR / 4;   // 16-byte chunks a thread copies per stage (the CTA has 2K threads, a stage 8K * SR / 16 chunks)
constexpr int NS = 3;    // stages in flight

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
    int dbg;   // timing probes (BPMF_BLOCK_DBG): 1 = no factorisation / solves, 2 = no Gram DMMAs, 4 = no solves; 0 = the product
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
};

// 1 / p for a positive normal p: MUFU.RCP64H seed + one cubic step (same as the K = 32 kernel)
__device__ __forceinline__ double fast_rcp(double p)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

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

template <int NB>
__global__ void __launch_bounds__(Cfg<NB>::T, (NB >= 12 ? 2 : NB >= 6 ? 4 : 8)) items_block_kernel(BlockArgs p)
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
        for (int a = tid; a < K; a += T) b[a] = rr0[a];
        // the K normals of this item: rng_set_pos((idx+1) * K * (iter+1)) (sample.cpp:266)
        if (warp == NWB - 1) warp_randn((uint32_t)(((long long)idx + 1) * (long long)K * ((long long)p.iter + 1)), K, z);
        __syncthreads();
        const int64_t ps = __ldg(p.colptr + idx), pe = __ldg(p.colptr + idx + 1);
        {
            // ---- gather + Gram of the item (computeMuLambda, sample.cpp:251-257)
            const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(smem);
            double *sw = reinterpret_cast<double *>(smem + C::W_OFF);
            double acc[NB + 1][2], r0 = 0.0, r1 = 0.0;
#pragma unroll
            for (int n = 0; n <= NB; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
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
            for (int st = 0; st < NS; ++st) { load_rows(st); issue(st); }
            load_rows(NS);
#pragma unroll 1
            for (int st = 0; st < nst; ++st) {
                cp_async_wait<NS - 1>();
                __syncthreads();
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
                __syncthreads();          // the slot is free
                issue(st + NS);
                load_rows(st + NS + 1);
            }
            cp_async_wait<0>();
            __syncthreads();              // the ring is dead: the tiles may overwrite it
            // MM = LambdaF + alpha * G (sample.cpp:297-298) into the tiles: lane (g, t) owns elements (g, 2t), (g, 2t + 1).
            // acc[n] is block (warp, n) for n <= warp and block (NB-1-warp, n-warp-1) after that.
#pragma unroll
            for (int n = 0; n <= NB; ++n) {
                const int I = (n <= warp) ? warp : NB - 1 - warp, J = (n <= warp) ? n : n - warp - 1;
                const int i = 8 * I + g, k = 8 * J + 2 * t;
                double2 v;
                v.x = fma(p.alpha, acc[n][0], __ldg(p.LambdaF + i + (size_t)k * K));
                v.y = fma(p.alpha, acc[n][1], __ldg(p.LambdaF + i + (size_t)(k + 1) * K));
                *reinterpret_cast<double2 *>(MMp + C::tile(I, J) + 8 * g + 2 * t) = v;
            }
            // rhs: sum over the quad's four ratings-of-a-group, on top of LambdaF * mu (sample.cpp:285) already in b
            r0 += __shfl_xor_sync(FULL, r0, 1); r0 += __shfl_xor_sync(FULL, r0, 2);
            r1 += __shfl_xor_sync(FULL, r1, 1); r1 += __shfl_xor_sync(FULL, r1, 2);
            if (t == 0) { b[8 * warp + g] += r0; b[8 * (NB - 1 - warp) + g] += r1; }
        }
        __syncthreads();
        if (p.dbg & 1) {
            if (warp == 0 && lane < 8) p.items[(size_t)idx * K + lane] = MMp[lane] + b[lane] + z[lane];
            continue;
        }
        // ---- chol.compute(MM) (sample.cpp:306) as MM = Lu D Lu^T, blocked by 8, on the tiles, with look-ahead: the
        // diagonal tile of block column kb + 1 is updated and factorised by warp 0 while the other warps run the rest of
        // the trailing update of block column kb, so its 8-pivot dependent chain is off the other warps' critical path.
        // factor_diag: one warp, the tile in registers (unscaled A~ on exit), lane k < 8 gets d_k and 1 / d_k.
        auto factor_diag = [&](double (&dg)[2], const int kb) {
            double myd = 1.0, myrinv = 1.0;
            double li[2] = {g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0};   // becomes Lu^-1 (row operations on I)
            bool ok = true;
#pragma unroll 1
            for (int k2 = 0; k2 < 4; ++k2) {
                const int qsrc = (lane & ~3) | k2;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 2 * k2 + e;
                    const double pv = __shfl_sync(FULL, dg[e], 4 * k + k2);          // pivot: lane (g = k, t = k2), register e
                    double bl0 = __shfl_sync(FULL, dg[e], 4 * (2 * t) + k2);         // a[2t][k], a[2t+1][k] (unscaled)
                    double bl1 = __shfl_sync(FULL, dg[e], 4 * (2 * t + 1) + k2);
                    const double ad = __shfl_sync(FULL, dg[e], qsrc);                // a[g][k]
                    const double wk0 = __shfl_sync(FULL, li[0], 4 * k + t), wk1 = __shfl_sync(FULL, li[1], 4 * k + t);   // row k of W
                    if (!(pv > 0.0)) ok = false;                                     // pivot <= 0 -> "Cholesky failed"
                    const double rinv = fast_rcp(pv);
                    if (lane == k) { myd = pv; myrinv = rinv; }
                    const double mg = (g > k) ? -(ad * rinv) : 0.0;                  // W[g][:] -= Lu(g,k) W[k][:]
                    li[0] = fma(mg, wk0, li[0]); li[1] = fma(mg, wk1, li[1]);
                    bl0 = (2 * t > k) ? -(bl0 * rinv) : 0.0;
                    bl1 = (2 * t + 1 > k) ? -(bl1 * rinv) : 0.0;
                    dg[0] = fma(ad, bl0, dg[0]); dg[1] = fma(ad, bl1, dg[1]);
                }
            }
            // the unit-lower factor divided by the pivots (Lu = A~ D^-1) goes back into the tile; d and 1/d into the vectors
            const double rv0 = __shfl_sync(FULL, myrinv, 2 * t), rv1 = __shfl_sync(FULL, myrinv, 2 * t + 1);
            *reinterpret_cast<double2 *>(MMp + C::tile(kb, kb) + 8 * g + 2 * t) = make_double2(dg[0] * rv0, dg[1] * rv1);
            *reinterpret_cast<double2 *>(MMp + C::LINV + 64 * kb + 8 * g + 2 * t) = make_double2(li[0], li[1]);
            if (lane < 8) { sd[8 * kb + lane] = myd; srinv[8 * kb + lane] = myrinv; }
            if (!ok && lane == 0) sint[1] = 1;
        };
        // the right-hand side as block row NB: row 0 of tile (NB, J) = b[8J .. 8J+7], rows 1..7 zero. Factorising the
        // matrix with this extra row leaves D^-1 Lu^-1 b in it (the forward solve, sample.cpp:321, for free).
        for (int e = tid; e < NB * 64; e += T) MMp[C::tile(NB, 0) + e] = ((e & 63) < 8) ? b[(e >> 6) * 8 + (e & 7)] : 0.0;
        if (warp == 0) {
            const double2 d2 = *reinterpret_cast<const double2 *>(MMp + C::tile(0, 0) + 8 * g + 2 * t);
            double dg[2] = {d2.x, d2.y};
            factor_diag(dg, 0);
        }
        __syncthreads();
#pragma unroll 1
        for (int kb = 0; kb < NB; ++kb) {
            // ---- panel: this warp's tiles below the (already factorised) diagonal tile, I = I0, I0 + NWB (block row NB is
            // the right-hand side):  Lu(I,kb) = A(I,kb) Lu(kb,kb)^-T D^-1, one 8x8x8 product on the tensor cores
            const int I0 = kb + 1 + ((warp - (kb + 1)) % NWB + NWB) % NWB;
            if (I0 <= NB) {
                const double *li = MMp + C::LINV + 64 * kb + 8 * g + t;     // B fragment: B[k][n] = Linv[n][k]
                const double lb0 = li[0], lb1 = li[4];
                const double rv0 = srinv[8 * kb + 2 * t], rv1 = srinv[8 * kb + 2 * t + 1];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int I = I0 + u * NWB;
                    if (I <= NB) {
                        double *tp = MMp + C::tile(I, kb);
                        double c0 = 0.0, c1 = 0.0;
                        dmma884(c0, c1, tp[8 * g + t], lb0);
                        dmma884(c0, c1, tp[8 * g + t + 4], lb1);
                        __syncwarp();                  // every lane has read the tile before it is overwritten
                        *reinterpret_cast<double2 *>(tp + 8 * g + 2 * t) = make_double2(c0 * rv0, c1 * rv1);
                    }
                }
            }
            __syncthreads();
            // ---- trailing tiles (I, J), kb < J <= I:  A(I,J) -= Lu(I,kb) D Lu(J,kb)^T  on the tensor cores
            if (kb + 1 < NB) {
                const double dc0 = sd[8 * kb + t], dc1 = sd[8 * kb + 4 + t];
                auto update_tile = [&](const int I, const int J) {
                    const double *ta = MMp + C::tile(I, kb) + 8 * g + t;
                    const double *tb = MMp + C::tile(J, kb) + 8 * g + t;
                    double2 cv = *reinterpret_cast<const double2 *>(MMp + C::tile(I, J) + 8 * g + 2 * t);
                    dmma884(cv.x, cv.y, -ta[0], tb[0] * dc0);
                    dmma884(cv.x, cv.y, -ta[4], tb[4] * dc1);
                    return cv;
                };
                // tile number e, row-major over the trailing triangle; e = 0 is the next diagonal tile
                constexpr int NOTH = NWB > 1 ? NWB - 1 : 1;       // warps that share the tiles e >= 1
                if (warp == 0) {
                    const double2 cv = update_tile(kb + 1, kb + 1);
                    double dg[2] = {cv.x, cv.y};
                    factor_diag(dg, kb + 1);
                }
                if (warp > 0 || NWB == 1) {
                    int I = kb + 1, J = kb + 1 + (NWB > 1 ? warp : 1);
                    while (I < NB && J > I) { J -= I - kb; ++I; }          // row I holds I - kb tiles (J = kb+1 .. I)
                    while (I < NB) {
                        const double2 cv = update_tile(I, J);
                        *reinterpret_cast<double2 *>(MMp + C::tile(I, J) + 8 * g + 2 * t) = cv;
                        J += NOTH;
                        while (I < NB && J > I) { J -= I - kb; ++I; }
                    }
                    for (; J < NB; J += NOTH) {       // block row NB (the right-hand side): tiles J = kb+1 .. NB-1
                        const double2 cv = update_tile(NB, J);
                        *reinterpret_cast<double2 *>(MMp + C::tile(NB, J) + 8 * g + 2 * t) = cv;
                    }
                }
            }
            __syncthreads();
        }
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
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
#pragma unroll 4
                for (int kk = 31; kk >= 0; --kk) {
                    const int k = 32 * r + kk;
                    if (k >= K) continue;
                    const double xk = __shfl_sync(FULL, v[r], kk);
                    const double *krow = MMp + ((k >> 3) * ((k >> 3) + 1) / 2) * 64 + 8 * (k & 7);   // Lu(k,i) = krow[colo(i)]
#pragma unroll
                    for (int r2 = 0; r2 <= r; ++r2) {
                        const int i = lane + 32 * r2;
                        if (i < k) v[r2] = fma(-krow[colo[r2]], xk, v[r2]);
                    }
                }
            }
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

template <int NB>
cudaError_t launch_nb(bpmf_gpu_ctx *c, const BlockArgs &p, long long n)
{
    using C = Cfg<NB>;
    static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
    auto kern = items_block_kernel<NB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, C::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    if (grid > n) grid = n;
    kern<<<(unsigned)grid, C::T, C::SMEM, c->stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

bool block_kernel_supports(int K) { return K % 16 == 0 && K >= 16 && K <= 128 && K != 32; }

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
    static const int dbg = [] { const char *v = getenv("BPMF_BLOCK_DBG"); return v ? atoi(v) : 0; }();
    p.dbg = dbg;
    cudaError_t e = cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned int), c->stream);
    if (e != cudaSuccess) return e;
    const long long n = (long long)s.to - s.from;
    if (n < 1) return cudaSuccess;
    switch (c->K) {
    case 16: e = launch_nb<2>(c, p, n); break;
    case 48: e = launch_nb<6>(c, p, n); break;
    case 64: e = launch_nb<8>(c, p, n); break;
    case 80: e = launch_nb<10>(c, p, n); break;
    case 96: e = launch_nb<12>(c, p, n); break;
    case 112: e = launch_nb<14>(c, p, n); break;
    case 128: e = launch_nb<16>(c, p, n); break;
    default: return cudaErrorInvalidValue;
    }
    c->launches++;
    return e;
}

}  // namespace bpmf
