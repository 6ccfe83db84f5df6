// stream_kernel.cu — the fused per-item conditional update for num_latent == 32 on sm_100a ("stream" variant):
// c++/sample.cpp:263-336 + 248-258 as ONE persistent kernel (items_stream32v3_kernel), one warp per item, 20 warps per SM.
//
//   gather   Every warp owns a ring of NS shared-memory stages of SR latent rows (256 B each, padded to 288 B).
//            The rows of the other factor named by the item's CSR entries are copied global -> shared with
//            cp.async (LDGSTS.128, L1-bypassing, zero-fill for the ragged end), NS-1 stages ahead of the
//            arithmetic and ACROSS item boundaries, so the HBM/L2 latency of the random 256-byte gathers is
//            overlapped with the tensor-core work and with the factorisation / solve tail of the previous item.
//            Consecutive items are claimed from a global counter by guided self-scheduling (remaining / resident warps items,
//            at most CLAIM, single items at the very end): their ratings are contiguous in the CSR arrays, so the <= 16 indices
//            and weights of a stage are one coalesced load, issued a whole stage ahead. The weights (v - mean) * alpha are
//            precomputed once per alpha (weights_kernel); a gather address is one IMAD.WIDE; lane 16 h + i holds the index of
//            the row its half-warp copies in step i.
//   Gram     fp64 tensor cores: mma.sync.m8n8k4 (DMMA). Lane 4g+t holds f[a] = y_t[8a+g] of rating t of a group of
//            four; f[I] is the A fragment and f[J] the B fragment of block (I,J), so ten DMMAs update the lower
//            triangle of 8x8 blocks and every gathered value is read from shared memory exactly once. The 288-byte
//            row stride makes those fragment loads bank-conflict free. rr += y * w is 4 DFMAs per group.
//   tail     tail32_warp: MM = LambdaF + alpha * G stays in the DMMA accumulator layout (20 registers per lane);
//            square-root-free blocked LDL^T (8x8 diagonal blocks with warp shuffles — the last two columns of a block peeled
//            off the loop —, trailing blocks with DMMAs whose B fragments are also the factor's panel blocks), the
//            unit-lower factor goes to shared memory divided by the pivots and negated, the two triangular solves are
//            shuffle + FMA chains in "lane j owns row j" form; K normals from Philox4x32-10 (rng.cuh).
//            What an instruction costs there, and what was tried: DESIGN.md section 5.
//   skew     items far heavier than the rest are cut into chunks (heavy_gram32_kernel / heavy_tail32_kernel below).
// Variants that were built and measured slower live in bench_micro/stream_experiments.cuh and bench_micro/stream_roles.cuh
// (compiled only with BPMF_STREAM_PROBES=1; logs in profiles/).
//
// Roofline: nnz * 256 B gathered per sweep against HBM; 2.5 DMMA (1280 flop) per rating against the fp64
// tensor pipe (37 TFLOP/s measured): at K = 32 the two are co-limiting (3.9 ms vs 3.5 ms for 100M ratings).
#include "common.cuh"
#include "rng.cuh"

#include <cuda.h>             // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <algorithm>
#include <cstdlib>

namespace bpmf {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SR = 16;                      // ratings (latent rows) per stage: four groups of four
constexpr int ROWB = 288;                   // bytes per staged row: 256 + 32 pad
constexpr int W_OFF = SR * ROWB;            // weights (SR doubles)
constexpr int META_OFF = W_OFF + SR * 8;    // int4 {n, item, first, last}
constexpr int MBAR_OFF = META_OFF + 16;      // one mbarrier per stage (bulk-copy variant)
constexpr int STAGE_BYTES = META_OFF + 32;
constexpr int LFS = 34;                     // row stride (doubles) of LambdaF in shared memory
constexpr int SHARED_BYTES = 32 * LFS * 8 + 32 * 8;
constexpr int TV_DEFAULT = 4 | 16 | 128 | 256;           // tail / fetch variant of the product kernels (see chol3_block_column, tail_variant)
constexpr int CLAIM = 16;                   // consecutive items claimed per atomic
constexpr int CLAIM_TAIL = 1;               // the smallest claim (the end of a sweep), so that all warps finish together (2 -> 1: 0.920 -> 0.910 ms on an eighth of Synthetic A)
// NS = stages per warp, NW = warps per CTA (one CTA per SM)
template <int NS> constexpr __host__ __device__ int warp_bytes() { return NS * STAGE_BYTES; }
static_assert(STAGE_BYTES % 16 == 0 && SHARED_BYTES % 16 == 0, "16-byte alignment for cp.async");

struct StreamArgs {
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
    const double *zero_row;   // 256 bytes of zeros in global memory (bulk-copy variant: source of the padding rows)
    int bulk_end;             // v3: items of [from, bulk_end) are claimed CLAIM at a time, [bulk_end, to) CLAIM_TAIL at a time
    int guided;               // > 0: guided self-scheduling (like the reference's schedule(guided), sample.cpp:352): a claim takes
                              // remaining / guided items, at most CLAIM, at least CLAIM_TAIL; bulk_end is not used
    int heavy_thr;            // SKIP: items with more ratings than this belong to the chunked path and are passed over
    const double *propLambda; // PROP: K*K x num per-item prior precisions (-m / -l, sample.cpp:272-277)
    int oob_row;              // gather4 variant: a row index outside the other side's latent matrix (zero fill)
    int sms;                  // (host side) SMs the kernel is launched on: item_sms()
    const double *wval;       // TV & 128: (val - mean_rating) * alpha per rating, precomputed (weights_kernel)
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// 16-byte global -> shared copy; src_bytes == 0 writes zeros without touching global memory
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier completion, used by the BULK variant of the v3 kernel
__device__ __forceinline__ void mbar_init(uint32_t mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, unsigned bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

constexpr __host__ __device__ int blk(int I, int J) { return I * (I + 1) / 2 + J; }

// one group of four staged ratings: 10 DMMAs on the lower triangle of blocks + the rhs
__device__ __forceinline__ void gram_group(double (&c)[10][2], double (&rrp)[4], const unsigned char *row, const double *wq)
{
    double f[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) f[a] = *reinterpret_cast<const double *>(row + a * 64);
    const double w = *wq;
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], f[I], f[J]);
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = fma(f[a], w, rrp[a]);
}

// =====================================================================================================================
// Version 3 of the same kernel. Two changes, both aimed at the fp64 pipe, which DMMA and scalar DFMA/DMUL share: a DMMA
// holds it for 16 cycles per scheduler, so every DEPENDENT scalar fp64 instruction of another warp's tail queues behind
// the Gram DMMAs, and the tail is one long dependent chain.
//   fetch   Index / value loads are aligned to the stage (one coalesced load of the next stage's <= 16 entries, issued a
//           whole stage ahead), so issuing a stage is 8 x (shuffle, address, cp.async) and little else.
//   tail    Square-root-free blocked LDL^T instead of LL^T: per column the chain is pivot shuffle -> reciprocal ->
//           multiply -> FMA (no rsqrt, no column scaling before the broadcasts); the unit-lower factor goes to shared
//           memory already divided by the pivots, which turns both triangular solves into pure shuffle + FMA chains:
//               A = Lu D Lu^T, L = Lu D^(1/2):  x = L^-T (L^-1 b + z) = Lu^-T (D^-1 Lu^-1 b + D^(-1/2) z)
//           One rsqrt per item (all 32 pivots at once, lane k owns d_k) instead of 32 serial ones.
//           bench_micro/emulate_block_ldlt.py is the lane-level model of chol3_block_column / the scatter / the solves.
// =====================================================================================================================
constexpr int LPACK1 = 496;                  // strictly lower triangle, column-major packed
constexpr int V3_ZY_OFF = LPACK1 * 8, V3_ZR_OFF = V3_ZY_OFF + 256, V3_B_OFF = V3_ZR_OFF + 256;
static_assert(V3_B_OFF + 256 <= STAGE_BYTES, "v3 tail scratch must fit in one stage");
constexpr __host__ __device__ int col_off1(int k) { return 31 * k - ((k * (k - 1)) / 2); }

// 1 / p for a positive normal p: MUFU.RCP64H seed (about 20 bits) + one cubic step. Relative error ~1e-17.
__device__ __forceinline__ double fast_rcp(double p)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// The column steps of one block column of the LDL^T. The last two columns of a block are peeled off the loop: column 7 has
// nothing to its right inside the block (only its pivot is recorded) and column 6 only column 7, so their zero-multiplier
// shuffles / multiplies / FMAs are not issued at all (7.55 -> 7.34 ms, profiles/r02b_tune_tail_variants.log).
// TV (tail variant) is a bit mask of measured alternatives, all bit-identical to TV == 0:
//   1  R1: every column step as a RANK-ONE DMMA with one-hot k-slots (config 15220). Lane (g, k2) already holds A~(I,KB)[g][k]
//      and D[g][k] (k = 2 k2 + e), so A fragment (t == k2 ? c[I][e] : 0) and B fragment (t == k2 && g > k ? -c[D][e] / d_k : 0)
//      need no lane exchange at all: per column 1 shuffle + (4 - KB) DMMAs instead of 3 + (4 - KB) shuffles and 2 (4 - KB)
//      DFMAs — 368 fewer fp64 / shuffle instructions per item (the three zero products add exactly), but 80 more DMMAs =
//      864 more cycles of the fp64 pipe per item: measured equal (profiles/r02_tune_rank1_dmma.log).
//   8  R1 for the block columns KB >= 2 only (config 19220): equal again (7.355 vs 7.338 ms)
//   2  pivots and their reciprocals go to shared memory (one 16-byte store per column by lane 0) instead of being picked up by
//      the lane that owns them with a compare + four selects per column; the lanes read them back when they need them
//   4  the panel blocks of the unit-lower factor are stored right after the trailing update's B fragments are formed: those
//      ARE -Lu (12 multiplies fewer per item); the factor is kept NEGATED in shared memory
template <int KB, int DBG = 0, int TV = 0>
__device__ __forceinline__ void chol3_block_column(double (&c)[10][2], double &myd, double &myrinv, int lane, int t, double2 *pr, double *Lp)
{
    constexpr int D = blk(KB, KB);
    constexpr bool R1 = (TV & 1) || ((TV & 8) && KB >= 2);
    constexpr bool PEEL = true, PR = (TV & 2) != 0;
    const int g = lane >> 2;
    const double p_probe = c[D][0] + 3.0;    // DBG & 64 (timing probe, wrong results): reciprocals that do not depend on the chain
    if (R1) {
#pragma unroll 1
        for (int k2 = 0; k2 < 4; ++k2) {
            const bool mine = (t == k2);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * k2 + e;
                const double p = __shfl_sync(FULL, c[D][e], 4 * k + k2);       // pivot d_k: lane (g = k, t = k2), register e
                const double rinv = fast_rcp((DBG & 64) ? p_probe + k : p);
                if (PR) { if (lane == 0) pr[8 * KB + k] = make_double2(p, rinv); }
                else if (lane == 8 * KB + k) { myd = p; myrinv = rinv; }
                if ((TV & 8) && k == 7) break;                                 // nothing right of column 7 (k2 == 3, e == 1: the last step)
                const double bf = (mine && g > k) ? -(c[D][e] * rinv) : 0.0;   // B[k2][n = g] = -D[n][k] / d_k below the pivot
#pragma unroll
                for (int I = KB; I < 4; ++I) {
                    const double af = mine ? c[blk(I, KB)][e] : 0.0;           // A[m = g][k2] = A~(I,KB)[m][k]
                    dmma884(c[blk(I, KB)][0], c[blk(I, KB)][1], af, bf);
                }
            }
        }
    } else {
#pragma unroll 1
    for (int k2 = 0; k2 < (PEEL ? 3 : 4); ++k2) {
        const int qsrc = (lane & ~3) | k2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = 2 * k2 + e;
            const double p = __shfl_sync(FULL, c[D][e], 4 * k + k2);       // pivot d_k: lane (g = k, t = k2), register e
            // a[2t][k], a[2t+1][k] (unscaled) and the column itself do not wait for the reciprocal
            double bl0 = __shfl_sync(FULL, c[D][e], 4 * (2 * t) + k2);
            double bl1 = __shfl_sync(FULL, c[D][e], 4 * (2 * t + 1) + k2);
            double a[4];
#pragma unroll
            for (int I = KB; I < 4; ++I) a[I] = __shfl_sync(FULL, c[blk(I, KB)][e], qsrc);   // a[8I+g][k]
            const double rinv = fast_rcp((DBG & 64) ? p_probe + k : p);
            if (PR) { if (lane == 0) pr[8 * KB + k] = make_double2(p, rinv); }
            else if (lane == 8 * KB + k) { myd = p; myrinv = rinv; }
            bl0 = (2 * t > k) ? -(bl0 * rinv) : 0.0;                        // zero where this lane's column is not right of k
            bl1 = (2 * t + 1 > k) ? -(bl1 * rinv) : 0.0;
#pragma unroll
            for (int I = KB; I < 4; ++I) {
                c[blk(I, KB)][0] = fma(a[I], bl0, c[blk(I, KB)][0]);
                c[blk(I, KB)][1] = fma(a[I], bl1, c[blk(I, KB)][1]);
            }
        }
    }
    if (PEEL) {
        // column 6 (k2 = 3, e = 0) updates column 7 only: lanes t == 3, register 1
        {
            const double p = __shfl_sync(FULL, c[D][0], 4 * 6 + 3);
            double bl1 = __shfl_sync(FULL, c[D][0], 4 * 7 + 3);              // D[7][6]
            double a[4];
#pragma unroll
            for (int I = KB; I < 4; ++I) a[I] = __shfl_sync(FULL, c[blk(I, KB)][0], lane | 3);   // a[8I+g][6]
            const double rinv = fast_rcp((DBG & 64) ? p_probe + 6 : p);
            if (PR) { if (lane == 0) pr[8 * KB + 6] = make_double2(p, rinv); }
            else if (lane == 8 * KB + 6) { myd = p; myrinv = rinv; }
            bl1 = (t == 3) ? -(bl1 * rinv) : 0.0;
#pragma unroll
            for (int I = KB; I < 4; ++I) c[blk(I, KB)][1] = fma(a[I], bl1, c[blk(I, KB)][1]);
        }
        // column 7: only its pivot
        {
            const double p = __shfl_sync(FULL, c[D][1], 4 * 7 + 3);
            const double rinv = fast_rcp((DBG & 64) ? p_probe + 7 : p);
            if (PR) { if (lane == 0) pr[8 * KB + 7] = make_double2(p, rinv); }
            else if (lane == 8 * KB + 7) { myd = p; myrinv = rinv; }
        }
    }
    }
    // trailing update A(I,J) -= A~(I,KB) D^-1 A~(J,KB)^T for KB < J <= I on the tensor cores
    // The sum over the block's eight columns may run in any order: DMMA number e takes column 2t + e in its k-slot t, so
    // the accumulator registers of A~(I,KB) and A~(J,KB) ARE the A and B fragments (lane 4g+t holds [g][2t + e]) and no
    // lane exchange is needed.
    if (PR) __syncwarp();                  // the block column's eight (pivot, reciprocal) pairs are in shared memory
    if (KB < 3) {
        double nrv[2], bs[4][2];
#pragma unroll
        for (int e = 0; e < 2; ++e) nrv[e] = PR ? -pr[8 * KB + 2 * t + e].y : -__shfl_sync(FULL, myrinv, 8 * KB + 2 * t + e);   // -1 / d of column 2t + e
#pragma unroll
        for (int J = KB + 1; J < 4; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) bs[J][e] = c[blk(J, KB)][e] * nrv[e];
        if (TV & 4) {
            // bs[J][e] = -Lu(8J + g, k) for column k = 8 KB + 2t + e: the factor's panel blocks, stored negated
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 8 * KB + 2 * t + e;
                double *lq = Lp + (31 * k - ((k * (k - 1)) >> 1)) + g - k - 1;
#pragma unroll
                for (int J = KB + 1; J < 4; ++J) lq[8 * J] = bs[J][e];
            }
        }
#pragma unroll
        for (int I = KB + 1; I < 4; ++I)
#pragma unroll
            for (int J = KB + 1; J <= I; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], c[blk(I, KB)][e], bs[J][e]);
    }
}

// The tail of one item for one warp (shared by the stream kernel and the heavy-item kernel): on entry the accumulators
// hold the item's Gram (DMMA layout) and rrp the quad-partial right-hand side; `stg` is TSCRATCH bytes of the warp's
// shared memory. Normals, MM = LambdaF + alpha G, LDL^T, solves, store (sample.cpp:266,297-324). Resets c and rrp.
// PROP: the item has its own prior precision (propagated posterior, sample.cpp:272-283): LambdaF is read from
// p.propLambda instead of shared memory, and rr starts from LambdaF_i * hp.mu with the GLOBAL hp.mu (quirk Q5), which the
// caller keeps in srr0.
template <int DBG, bool PROP = false, int TV = 0>
__device__ __forceinline__ void tail32_warp(double (&c)[10][2], double (&rrp)[4], const int idx, unsigned char *stg, const double *sLF,
                                            const double *srr0, const StreamArgs &p, const int lane, unsigned char *vecs = nullptr)
{
    if (!vecs) vecs = stg + V3_ZY_OFF;    // zy | zr | wb, 256 bytes each; by default behind the packed factor
    const double *LF = PROP ? p.propLambda + (size_t)idx * 1024 : sLF;   // LambdaF(i,k) at LF[k * LS + i]
    constexpr int LS = PROP ? 32 : LFS;
    double rr0 = 0.0;
    if (PROP) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) rr0 = fma(__ldg(LF + j * 32 + lane), srr0[j], rr0);   // rr = hp_LambdaF * hp.mu (sample.cpp:285)
    }
    const int g = lane >> 2, t = lane & 3;
    double *zy = reinterpret_cast<double *>(vecs), *zr = reinterpret_cast<double *>(vecs + 256);
    double *wb = reinterpret_cast<double *>(vecs + 512), *Lp = reinterpret_cast<double *>(stg);
    double2 *prv = reinterpret_cast<double2 *>(vecs);     // TV & 2: (d_k, 1 / d_k) pairs, over zy | zr once z has been formed
    constexpr bool PR = (TV & 2) != 0, NEGL = (TV & 4) != 0;
    // the K normals of this item: rng_set_pos((idx+1)*K*(iter+1)) (sample.cpp:266). Accepted polar attempts are numbered
    // by ballot; lane n then finishes normal n (one log / sqrt / divide per lane instead of one per attempt).
    if (!(DBG & 8)) {
        const uint32_t seed = (uint32_t)(((long long)idx + 1) * 32ll * ((long long)p.iter + 1));
        int have = 0;
        for (uint32_t base = 0; have < 32; base += 32) {
            const U4 bk = stream_block(seed, base + lane);
            const Polar pa = polar_attempt(bk.v[3], bk.v[2], bk.v[1], bk.v[0]);
            const unsigned m = __ballot_sync(FULL, pa.ok);
            const int n = have + __popc(m & ((1u << lane) - 1u));
            if (pa.ok && n < 32) { zy[n] = pa.y; zr[n] = pa.r2; }
            have += __popc(m);
        }
    }
    // rr = LambdaF*mu + sum over the quad's four ratings-of-a-group
    double bb;
    if (TV & 256) {
        // through shared memory: lane 4g + t puts its four partials where lane r = 8a + g finds the quad's four values of ITS
        // row side by side, and every lane adds its own row in the order of the shuffle tree, (x0 + x1) + (x2 + x3): 4 stores,
        // 2 loads and 4 additions instead of 16 shuffles, 12 additions, 4 stores and a load
        double *rs = Lp;
#pragma unroll
        for (int a = 0; a < 4; ++a) rs[a * 32 + lane] = rrp[a];
        __syncwarp();
        const double2 v01 = *reinterpret_cast<const double2 *>(rs + 4 * lane), v23 = *reinterpret_cast<const double2 *>(rs + 4 * lane + 2);
        bb = (PROP ? 0.0 : srr0[lane]) + ((v01.x + v01.y) + (v23.x + v23.y));
        __syncwarp();                     // (the factor is written over rs later)
    } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        rrp[a] += __shfl_xor_sync(FULL, rrp[a], 1);
        rrp[a] += __shfl_xor_sync(FULL, rrp[a], 2);
    }
    if (t == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) wb[8 * a + g] = (PROP ? 0.0 : srr0[8 * a + g]) + rrp[a];
    }
    __syncwarp();
    bb = wb[lane];
    }
    const double z = (DBG & 8) ? 0.25 * lane : __dmul_rn(zy[lane], polar_mult(zr[lane]));
    if (PROP) bb += rr0;
    if (PR) __syncwarp();                 // zy | zr are free: the pivots go there
    // MM = LambdaF + alpha * G (sample.cpp:297-298), in place in the accumulator layout
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                c[blk(I, J)][e] = fma(p.alpha, c[blk(I, J)][e], LF[(8 * J + 2 * t + e) * LS + 8 * I + g]);   // LambdaF(i,k), i >= k
    // chol.compute(MM) (sample.cpp:306) as MM = Lu D Lu^T; lane k ends up with d_k and 1 / d_k
    double myd = 1.0, myrinv = 1.0;
    bool ok = true;
    if (!(DBG & 16)) {
        chol3_block_column<0, DBG, TV>(c, myd, myrinv, lane, t, prv, Lp);
        chol3_block_column<1, DBG, TV>(c, myd, myrinv, lane, t, prv, Lp);
        chol3_block_column<2, DBG, TV>(c, myd, myrinv, lane, t, prv, Lp);
        chol3_block_column<3, DBG, TV>(c, myd, myrinv, lane, t, prv, Lp);
        if (PR) { const double2 m = prv[lane]; myd = m.x; myrinv = m.y; }
    } else {
        myd = c[0][0] + 2.0; myrinv = fast_rcp(myd);
        if (PR) { prv[lane] = make_double2(myd, myrinv); __syncwarp(); }
    }
    // Eigen LLT: a pivot <= 0 -> "Cholesky failed". Lane k holds d_k; a bad pivot poisons what follows (NaN also fails)
    ok = __all_sync(FULL, myd > 0.0) || (DBG & 64);   // (the probe's factorization is garbage: keep the solves in the timing)
    const double myrs = rsqrt(myd);                       // 1 / L(k,k)
    // Lu -> shared memory, packed by columns without the unit diagonal: element (i,k), i > k, at col_off1(k) + i - k - 1
    {
#pragma unroll
        for (int J = 0; J < 4; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 8 * J + 2 * t + e;
                const double rk = PR ? prv[k].y : __shfl_sync(FULL, myrinv, k);
                double *lq = Lp + (31 * k - ((k * (k - 1)) >> 1)) + g - k - 1;
                if (NEGL) {               // the panel blocks are in place already (chol3_block_column), negated like these
                    if (g > 2 * t + e) lq[8 * J] = c[blk(J, J)][e] * -rk;
                } else {
#pragma unroll
                for (int I = J; I < 4; ++I)
                    if (I > J || g > 2 * t + e) lq[8 * I] = c[blk(I, J)][e] * rk;
                }
            }
    }
    // reset the accumulators for the next item
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;
    __syncwarp();
    if (ok) {
        // chol.matrixL().solveInPlace(rr); rr += nrandn(); chol.matrixU().solveInPlace(rr) (sample.cpp:321-323):
        // lane j owns row j; one broadcast + one FMA per step.
        if (!(DBG & 32)) {
            const double *lf = Lp + lane - 1;              // element (lane, k) at lf[col_off1(k) - k]
#pragma unroll
            for (int k = 0; k < 31; ++k) {
                const double yk = __shfl_sync(FULL, bb, k);
                if (lane > k) bb = NEGL ? fma(lf[col_off1(k) - k], yk, bb) : fma(-lf[col_off1(k) - k], yk, bb);
            }
        }
        double yv = fma(bb, myrinv, myrs * z);             // D^-1 Lu^-1 b + D^(-1/2) z
        if (!(DBG & 32)) {
            const double *lb = Lp + (31 * lane - ((lane * (lane - 1)) >> 1)) - lane - 1;   // element (i, lane) at lb[i]
#pragma unroll
            for (int i = 31; i >= 1; --i) {
                const double xi = __shfl_sync(FULL, yv, i);
                if (lane < i) yv = NEGL ? fma(lb[i], xi, yv) : fma(-lb[i], xi, yv);
            }
        }
        // items().col(idx) = rr (sample.cpp:324); push to the peer replicas (replaces send_item, :370)
        p.items[(size_t)idx * 32 + lane] = yv;
        for (int q = 0; q < p.npeers; ++q) {
            double *dst = p.peers[q];
            if (dst && dst != p.items) dst[(size_t)idx * 32 + lane] = yv;
        }
    } else if (lane == 0) {           // THROWERROR("Cholesky failed") (sample.cpp:308): reported through the error word
        atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
    }
}

// DBG (bench_micro/tune_stream.py only) is a bit mask: 0 = the product; 1 = no tail (Gram only); 2 = no Gram DMMAs;
// 4 = no gather (Gram on whatever the stage holds); 8 = no normals; 16 = no factorization; 32 = no triangular solves.
// Anything but 0 produces garbage: the probes exist to time the parts of the kernel.
// BULK: the gather uses one TMA bulk copy (cp.async.bulk, 256 B) per latent row, issued by the lane that holds the row's
// index and completed on the stage's mbarrier, instead of sixteen 16-byte cp.async per row spread over half a warp.
// TOK: "Gram token". DMMA and scalar fp64 share one pipe per scheduler and the arbiter is not fair to the scalar side:
// with two warps of a scheduler streaming DMMAs a dependent DFMA chain of a third warp runs at 56 cycles per instruction
// instead of 8 (14.6 with one streaming warp; profiles/r01_fp64_latency_microbench.txt), which is what made the tails
// slow. With TOK != 0 at most ONE warp per scheduler (warp id % 4) executes Gram DMMAs at a time: 1 = the token is held
// per stage (<= 40 DMMAs), 2 = per item. A single warp saturates the DMMA pipe, so the Gram loses nothing.
// TOK == 3 ("phased"): the warps of a CTA run their Gram phases together and their tails together (two CTA barriers
// per item). Nobody is in a tail while DMMAs stream, so the tails' dependent scalar fp64 chains run at the uncontended
// 8 cycles per instruction instead of 15-56, and the Gram phase keeps the DMMA pipe full with every warp.
// SKIP: the range holds heavy items (more than p.heavy_thr ratings). They are sampled by heavy_gram32_kernel /
// heavy_tail32_kernel, which run AFTER this kernel on the same stream; here they are gathered as if they had no ratings
// (a prior-only draw that the heavy path overwrites), which keeps the fetch state machine contiguous.
template <int NS, int NW, int DBG, bool BULK, int TOK, bool SKIP = false, bool PROP = false, int TV = 0>
__global__ void __launch_bounds__(NW * 32, 1) items_stream32v3_kernel(StreamArgs p)
{
    constexpr int WARP_BYTES = warp_bytes<NS>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    unsigned char *wbase = smem_raw + SHARED_BYTES + (size_t)warp * WARP_BYTES;
    const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);

    __shared__ int s_token[4];            // TOK: 0 = free, 1 = a warp of that scheduler is in its Gram burst
    if (tid < 4) s_token[tid] = 0;
    int *my_token = &s_token[warp & 3];
    bool have_token = false;
    for (int e = tid; e < 1024; e += NW * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = PROP ? p.mu[tid] : s;                                 // PROP: per-item LambdaF, the tail needs hp.mu itself
    }
    __syncthreads();

    // ---------------- fetch-side state (warp-uniform unless noted); positions are relative to the group's first rating
    int g_base = 0, g_n = 0, f_it = 0;
    int cpr = 0;                          // per lane: colptr[g_base + lane] - colptr[g_base]
    int f_pos = 0, f_end = 0, f_start = 0, g_end = 0;
    constexpr bool PW = (TV & 128) != 0 && !BULK;   // the ratings' weights are precomputed: p.wval instead of p.val
    const int32_t *g_idx = p.rowidx;      // rowidx / val at the group's first rating
    const double *g_val = PW ? p.wval : p.val;
    int32_t n_idx = 0;                    // per lane: index / value of stream position f_pos + (lane & 15): the NEXT stage's
    double n_w = 0.0;
    bool f_done = false;
    const unsigned char *src_lane = reinterpret_cast<const unsigned char *>(p.other) + (lane & 15) * 16;   // this lane's 16 B of a row
    const uint32_t dst_lane = BULK ? wbase_s + (lane & 15) * ROWB : wbase_s + (lane >> 4) * ROWB + (lane & 15) * 16;
    const int half = lane >> 4;
    unsigned phases = 0;                  // BULK: bit h = parity of the phase of stage h the consumer waits for next
    if (BULK) {
        if (lane == 0) {
#pragma unroll 1
            for (int s = 0; s < NS; ++s) mbar_init(wbase_s + s * STAGE_BYTES + MBAR_OFF, 1);
        }
        fence_proxy_async();
        __syncwarp();
    }

    // TV & 16: lane 16 h + i (i < 8) holds the index / value of the stage's row 2 i + h, the row its half-warp copies in step i,
    // so that the shuffle of step i names its source lane by an immediate (width 16) instead of a per-lane register
    constexpr bool FPERM = (TV & 16) != 0 && !BULK;
    const int rowl = FPERM ? 2 * (lane & 7) + (lane >> 4) : (lane & 15);
    auto load_next = [&]() {
        const int q = f_pos + rowl;
        n_idx = 0; n_w = 0.0;
        if (q < g_end) {                  // entries past the item's end but inside the group are loaded and never used
            n_idx = __ldg(g_idx + q);
            n_w = __ldg(g_val + q);
        }
    };
    auto claim = [&]() {
        int base = 0, lim = p.bulk_end;
        if (lane == 0) {
            if (p.guided > 0) {
                // the size comes from a slightly stale read of the counter; whatever it is, [base, base + sz) is this warp's alone
                const int seen = (int)*reinterpret_cast<volatile unsigned int *>(p.work_counter);
                const int sz = max(CLAIM_TAIL, min(CLAIM, (p.to - p.from - seen) / p.guided));
                base = p.from + (int)atomicAdd(p.work_counter, (unsigned)sz);
                lim = min(p.to, base + sz);
            } else {
                base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
                if (base >= p.bulk_end) {     // small groups at the end of the sweep
                    base = p.bulk_end + (int)atomicAdd(p.work_counter + 1, (unsigned)CLAIM_TAIL);
                    lim = min(p.to, base + CLAIM_TAIL);
                }
            }
        }
        base = __shfl_sync(FULL, base, 0);
        lim = __shfl_sync(FULL, lim, 0);
        if (base >= p.to) { f_done = true; return; }
        g_base = base;
        g_n = min(CLAIM, lim - base);
        const int64_t c0 = __ldg(p.colptr + base);
        cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
        g_idx = p.rowidx + c0;
        g_val = (PW ? p.wval : p.val) + c0;
        g_end = __shfl_sync(FULL, cpr, g_n);
        f_it = 0;
        f_start = 0;
        f_end = __shfl_sync(FULL, cpr, 1);
        if (SKIP && f_end - f_start > p.heavy_thr) f_start = f_end;     // heavy: gathered as if it had no ratings
        f_pos = f_start;
    };
    // fill ring slot `slot` with the next (at most SR) ratings of the current item; exactly one commit_group per call
    auto issue_stage = [&](int slot) {
        const uint32_t st = dst_lane + slot * STAGE_BYTES;
        unsigned char *stg = wbase + slot * STAGE_BYTES;
        if (f_done) {
            if (lane == 0) {
                *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(-1, 0, 0, 0);
                if (BULK) mbar_arrive_expect_tx(wbase_s + slot * STAGE_BYTES + MBAR_OFF, 0);
            }
            if (!BULK) cp_async_commit();
            return;
        }
        const int n = min(SR, f_end - f_pos);
        const int nn = n - half;
        if (BULK) {
            // lane r < 16 owns row r of the stage: one 256-byte bulk copy from the row its own index names; the rows that
            // pad the last group of four come from a row of zeros. The stage was last touched through the generic proxy
            // (fragment loads, tail scratch), so order those before the async-proxy writes.
            const uint32_t mbar = wbase_s + slot * STAGE_BYTES + MBAR_OFF;
            const int n4 = (n + 3) & ~3;
            fence_proxy_async();
            if (lane == 0) mbar_arrive_expect_tx(mbar, (unsigned)n4 * 256u);
            __syncwarp();
            if (lane < n4 && !(DBG & 4)) {
                const void *src = (lane < n) ? static_cast<const void *>(p.other + (size_t)(unsigned)n_idx * 32) : static_cast<const void *>(p.zero_row);
                bulk_g2s(st, src, 256u, mbar);
            } else if (lane < n4) {
                bulk_g2s(st, p.zero_row, 256u, mbar);
            }
        } else {
            if (FPERM && n == SR) {       // a full stage: no predicates
#pragma unroll
                for (int i = 0; i < SR / 2; ++i) {
                    const unsigned j = (unsigned)__shfl_sync(FULL, n_idx, i, 16);
                    unsigned long long src;
                    asm("mad.wide.u32 %0, %1, 256, %2;" : "=l"(src) : "r"(j), "l"(src_lane));
                    if (!(DBG & 4)) cp_async16(st + 2 * i * ROWB, reinterpret_cast<const void *>(src), 16);
                }
            } else
#pragma unroll
            for (int i = 0; i < SR / 2; ++i) {
                const unsigned j = (unsigned)(FPERM ? __shfl_sync(FULL, n_idx, i, 16) : __shfl_sync(FULL, n_idx, 2 * i + half));
                // rows past the item's end are zero-filled (src-size 0 reads nothing; j is still a valid row)
                // one 64-bit multiply-add per source address (IMAD.WIDE.U32 with the lane's base as the addend; written out,
                // the compiler makes it four instructions: 7.69 -> 7.55 ms)
                unsigned long long src;
                asm("mad.wide.u32 %0, %1, 256, %2;" : "=l"(src) : "r"(j), "l"(src_lane));
                if (!(DBG & 4)) cp_async16(st + 2 * i * ROWB, reinterpret_cast<const void *>(src), (2 * i < nn) ? 16 : 0);
            }
        }
        // rr weight (v - mean_rating) * alpha (sample.cpp:255)
        // (PW: rows past the item's end may carry the next item's weight; their latent rows are zero-filled, so they add 0)
        if (FPERM ? (lane & 8) == 0 : lane < SR)
            reinterpret_cast<double *>(stg + W_OFF)[rowl] = PW ? n_w : (rowl < n) ? (n_w - p.mean_rating) * p.alpha : 0.0;
        const int last = (f_pos + n == f_end);
        if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(n, g_base + f_it, f_pos == f_start, last);
        if (!BULK) cp_async_commit();
        f_pos += n;
        if (last) {
            ++f_it;
            if (f_it >= g_n) claim();
            else {
                f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1);
                if (SKIP && f_end - f_start > p.heavy_thr) { f_start = f_end; f_pos = f_start; }
            }
        }
        if (!f_done) load_next();
    };

    claim();
    if (!f_done) load_next();
#pragma unroll 1
    for (int s = 0; s < NS; ++s) issue_stage(s);

    double c[10][2];
    double rrp[4];
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;

    int h = 0;
#pragma unroll 1
    for (;;) {
        if (BULK) {
            mbar_wait(wbase_s + h * STAGE_BYTES + MBAR_OFF, (phases >> h) & 1u);
            phases ^= 1u << h;
        } else {
            cp_async_wait<NS - 1>();
        }
        __syncwarp();
        unsigned char *stg = wbase + h * STAGE_BYTES;
        const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
        if (meta.x < 0) {
            if (TOK == 3) {               // out of items: keep the CTA's barrier protocol going until every warp is
                for (;;) {
                    if (!__syncthreads_or(0)) break;
                    __syncthreads();
                }
            }
            break;
        }
        // ---------------- Gram + rhs of this stage (computeMuLambda, sample.cpp:251-257) ----------------
        {
            const unsigned char *row = stg + t * ROWB + g * 8;
            const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
            if (!(DBG & 2)) {
                if (TOK && meta.x > 0 && !have_token) {
                    if (lane == 0) while (atomicCAS(my_token, 0, 1) != 0) __nanosleep(20);
                    __syncwarp();
                    have_token = true;
                }
                if (meta.x == SR) {       // full stage: straight-line, so the fragment loads run ahead of the DMMAs
                    gram_group(c, rrp, row, wq);
                    gram_group(c, rrp, row + 4 * ROWB, wq + 4);
                    gram_group(c, rrp, row + 8 * ROWB, wq + 8);
                    gram_group(c, rrp, row + 12 * ROWB, wq + 12);
                } else {
                    if (meta.x > 0) gram_group(c, rrp, row, wq);
                    if (meta.x > 4) gram_group(c, rrp, row + 4 * ROWB, wq + 4);
                    if (meta.x > 8) gram_group(c, rrp, row + 8 * ROWB, wq + 8);
                    if (meta.x > 12) gram_group(c, rrp, row + 12 * ROWB, wq + 12);
                }
                if (TOK && have_token && (TOK == 1 || meta.w)) {
                    __syncwarp();
                    if (lane == 0) atomicExch(my_token, 0);
                    have_token = false;
                }
            } else if (meta.x > 0) {
                c[0][0] += *reinterpret_cast<const double *>(row) * *wq;
            }
        }
        __syncwarp();                     // every lane is done reading slot h
        if (!meta.w) {                    // more stages of this item to come: refill the slot and go on
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
            continue;
        }
        if (TOK == 3) __syncthreads_or(1);   // phase boundary: every warp's Gram is complete
        // ---------------- tail: one item's Gram is complete; slot h is its scratch until the refill at the end ---------
        const int idx = meta.y;
        if (DBG & 1) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 10; ++b) { acc += c[b][0] + c[b][1]; c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
            for (int a = 0; a < 4; ++a) { acc += rrp[a]; rrp[a] = 0.0; }
            p.items[(size_t)idx * 32 + lane] = acc;
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
            continue;
        }
        tail32_warp<DBG, PROP, TV>(c, rrp, idx, stg, sLF, srr0, p, lane);
        __syncwarp();                     // the scratch is free again
        issue_stage(h);
        h = (h + 1 == NS) ? 0 : h + 1;
        if (TOK == 3) __syncthreads();    // phase boundary: every warp's tail is complete
    }
    if (!BULK) cp_async_wait<0>();
}

// =====================================================================================================================
// The v3 kernel with the gather on the TMA unit in its Blackwell GATHER form: cp.async.bulk.tensor.2d ... tile::gather4
// (SASS UTMALDG) copies FOUR index-addressed rows of the other side's latent matrix per instruction, through a 2-D tensor
// map of that matrix (box = 16 doubles x 1 row, SWIZZLE_128B; a latent row is two boxes), completion counted in bytes on
// the stage's mbarrier. A stage of 16 ratings is 8 such copies, issued by lanes 0..7 in one instruction, instead of
// 8 x (shuffle + address + cp.async) by the whole warp.
// Shared-memory layout of a stage: the LEFT halves (latent 0..15) of its 16 rows are 16 lines of 128 bytes (two swizzle
// atoms of 1024 bytes), the RIGHT halves another 16 lines. TMA cannot pad rows, so bank conflicts of the fragment loads
// are avoided by WHERE a rating goes instead: rating 4q + t (k-slot t of DMMA group q) sits in line 8 (q >> 1) + 2 t + (q & 1),
// so the four ratings of a group are in lines of equal parity and the 128-byte swizzle (16-byte chunk index XOR line
// index) spreads a half-warp's 8 x 32 bytes over all 32 banks. Rating slots past the item's end name row `num_other`,
// which is outside the tensor: the TMA unit fills them with zeros.
// =====================================================================================================================
constexpr int G4_ROWS_BYTES = SR * 256;                      // per stage, 1024-byte aligned
constexpr int G4_AUX_STAGE = SR * 8 + 32;                    // weights | meta int4 | mbarrier
constexpr int G4_VECS = 768;                                 // per warp: zy | zr | wb of the tail
template <int NS> constexpr __host__ __device__ int g4_aux_warp() { return NS * G4_AUX_STAGE + G4_VECS; }
constexpr int G4_SHARED = ((SHARED_BYTES + 1023) / 1024) * 1024;
static_assert(LPACK1 * 8 <= G4_ROWS_BYTES, "the packed factor must fit in a stage's rows");

__device__ __forceinline__ void tma_gather4(uint32_t dst, const void *tmap, int col, int r0, int r1, int r2, int r3, uint32_t mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                 ::"r"(dst), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(mbar)
                 : "memory");
}

// one group of four staged ratings in the gather4 layout: lo = left-half line of this lane's rating, already including
// the lane's 8-byte slot; sw0 / sw1 = the swizzled 16-byte chunk offsets for fragment pairs (0, 2) and (1, 3)
__device__ __forceinline__ void gram_group_g4(double (&c)[10][2], double (&rrp)[4], const unsigned char *line, int sw0, int sw1, const double *wq)
{
    double f[4];
    f[0] = *reinterpret_cast<const double *>(line + sw0);
    f[1] = *reinterpret_cast<const double *>(line + sw1);
    f[2] = *reinterpret_cast<const double *>(line + 2048 + sw0);
    f[3] = *reinterpret_cast<const double *>(line + 2048 + sw1);
    const double w = *wq;
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], f[I], f[J]);
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = fma(f[a], w, rrp[a]);
}

template <int NS, int NW, bool SKIP>
__global__ void __launch_bounds__(NW * 32, 1) items_stream32g4_kernel(StreamArgs p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) unsigned char smem_g4[];
    unsigned char *const smem_raw = smem_g4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    unsigned char *rows = smem_raw + G4_SHARED + (size_t)warp * NS * G4_ROWS_BYTES;
    unsigned char *aux = smem_raw + G4_SHARED + (size_t)NW * NS * G4_ROWS_BYTES + (size_t)warp * g4_aux_warp<NS>();
    unsigned char *vecs = aux + NS * G4_AUX_STAGE;
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows), aux_s = (uint32_t)__cvta_generic_to_shared(aux);

    if ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u) {   // the 128-byte swizzle pattern follows the address bits
        if (tid == 0) atomicMax(p.err, ERR_CHOLESKY | 0xfffffffeull);
        return;
    }
    for (int e = tid; e < 1024; e += NW * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < NS; ++s) mbar_init(aux_s + s * G4_AUX_STAGE + SR * 8 + 16, 1);
    }
    fence_proxy_async();
    __syncthreads();

    // ---------------- fetch-side state (warp-uniform unless noted)
    int g_base = 0, g_n = 0, f_it = 0;
    int cpr = 0;                          // per lane: colptr[g_base + lane] - colptr[g_base]
    int f_pos = 0, f_end = 0, f_start = 0, g_end = 0;
    const int32_t *g_idx = p.rowidx;
    const double *g_val = p.val;
    int32_t n_idx = 0;                    // per lane: index / value of stream position f_pos + (lane & 15): the NEXT stage's
    double n_w = 0.0;
    bool f_done = false;
    unsigned phases = 0;                  // bit h = parity of the phase of stage h the consumer waits for next
    // this lane's copy (lanes 0..7): atom A = groups 2A and 2A + 1, lines 4h .. 4h + 3 of the atom, half hf of the latent row;
    // it gathers ratings 8A + 2h + {0, 4, 1, 5} (see the layout above)
    const int cA = (lane >> 2) & 1, ch = (lane >> 1) & 1, chf = lane & 1;
    const int cr0 = 8 * cA + 2 * ch;      // ratings cr0, cr0 + 4, cr0 + 1, cr0 + 5
    const uint32_t cdst = rows_s + chf * 2048 + (8 * cA + 4 * ch) * 128;
    const int oob = p.oob_row;            // a row outside the tensor: zero fill

    auto load_next = [&]() {
        const int q = f_pos + (lane & 15);
        n_idx = oob; n_w = 0.0;
        if (q < min(g_end, f_end)) {
            n_idx = __ldg(g_idx + q);
            n_w = __ldg(g_val + q);
        }
    };
    auto claim = [&]() {
        int base = 0, lim = p.bulk_end;
        if (lane == 0) {
            if (p.guided > 0) {
                const int seen = (int)*reinterpret_cast<volatile unsigned int *>(p.work_counter);
                const int sz = max(CLAIM_TAIL, min(CLAIM, (p.to - p.from - seen) / p.guided));
                base = p.from + (int)atomicAdd(p.work_counter, (unsigned)sz);
                lim = min(p.to, base + sz);
            } else {
                base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
                if (base >= p.bulk_end) {
                    base = p.bulk_end + (int)atomicAdd(p.work_counter + 1, (unsigned)CLAIM_TAIL);
                    lim = min(p.to, base + CLAIM_TAIL);
                }
            }
        }
        base = __shfl_sync(FULL, base, 0);
        lim = __shfl_sync(FULL, lim, 0);
        if (base >= p.to) { f_done = true; return; }
        g_base = base;
        g_n = min(CLAIM, lim - base);
        const int64_t c0 = __ldg(p.colptr + base);
        cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
        g_idx = p.rowidx + c0;
        g_val = p.val + c0;
        g_end = __shfl_sync(FULL, cpr, g_n);
        f_it = 0;
        f_start = 0;
        f_end = __shfl_sync(FULL, cpr, 1);
        if (SKIP && f_end - f_start > p.heavy_thr) f_start = f_end;     // heavy: gathered as if it had no ratings
        f_pos = f_start;
    };
    // fill ring slot `slot` with the next (at most SR) ratings of the current item
    auto issue_stage = [&](int slot) {
        unsigned char *ax = aux + slot * G4_AUX_STAGE;
        const uint32_t mbar = aux_s + slot * G4_AUX_STAGE + SR * 8 + 16;
        if (f_done) {
            if (lane == 0) {
                *reinterpret_cast<int4 *>(ax + SR * 8) = make_int4(-1, 0, 0, 0);
                mbar_arrive_expect_tx(mbar, 0);
            }
            return;
        }
        const int n = min(SR, f_end - f_pos);
        // the four rows of this lane's copy; n_idx already holds the out-of-tensor row for slots past the item's end
        const int r0 = __shfl_sync(FULL, n_idx, cr0), r1 = __shfl_sync(FULL, n_idx, cr0 + 4);
        const int r2 = __shfl_sync(FULL, n_idx, cr0 + 1), r3 = __shfl_sync(FULL, n_idx, cr0 + 5);
        const int nops = n > 8 ? 8 : (n > 0 ? 4 : 0);               // ratings 8..15 live in the second atom
        // the stage was last touched through the generic proxy (fragment loads, tail scratch): order those before the TMA writes
        fence_proxy_async();
        if (lane == 0) mbar_arrive_expect_tx(mbar, (unsigned)nops * 512u);
        __syncwarp();
        if (lane < nops) tma_gather4(cdst + slot * G4_ROWS_BYTES, &tmap, 16 * chf, r0, r1, r2, r3, mbar);
        // rr weight (v - mean_rating) * alpha (sample.cpp:255)
        if (lane < SR) reinterpret_cast<double *>(ax)[lane] = (lane < n) ? (n_w - p.mean_rating) * p.alpha : 0.0;
        const int last = (f_pos + n == f_end);
        if (lane == 0) *reinterpret_cast<int4 *>(ax + SR * 8) = make_int4(n, g_base + f_it, f_pos == f_start, last);
        f_pos += n;
        if (last) {
            ++f_it;
            if (f_it >= g_n) claim();
            else {
                f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1);
                if (SKIP && f_end - f_start > p.heavy_thr) { f_start = f_end; f_pos = f_start; }
            }
        }
        if (!f_done) load_next();
    };

    claim();
    if (!f_done) load_next();
#pragma unroll 1
    for (int s = 0; s < NS; ++s) issue_stage(s);

    double c[10][2];
    double rrp[4];
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;

    // fragment addressing of this lane: group q's rating t is in line 8 (q >> 1) + 2 t + (q & 1); chunk index XOR (line & 7)
    const int y0 = (g >> 1) ^ (2 * t);                             // groups 0, 2 (even lines)
    const int e_sw0 = (y0 << 4) + (g & 1) * 8, e_sw1 = ((y0 ^ 4) << 4) + (g & 1) * 8;
    const int o_sw0 = ((y0 ^ 1) << 4) + (g & 1) * 8, o_sw1 = ((y0 ^ 5) << 4) + (g & 1) * 8;
    const int l_even = 2 * t * 128, l_odd = (2 * t + 1) * 128;

    int h = 0;
#pragma unroll 1
    for (;;) {
        mbar_wait(aux_s + h * G4_AUX_STAGE + SR * 8 + 16, (phases >> h) & 1u);
        phases ^= 1u << h;
        unsigned char *st = rows + h * G4_ROWS_BYTES;
        unsigned char *ax = aux + h * G4_AUX_STAGE;
        const int4 meta = *reinterpret_cast<const int4 *>(ax + SR * 8);
        if (meta.x < 0) break;
        {
            const double *wq = reinterpret_cast<const double *>(ax) + t;
            if (meta.x == SR) {
                gram_group_g4(c, rrp, st + l_even, e_sw0, e_sw1, wq);
                gram_group_g4(c, rrp, st + l_odd, o_sw0, o_sw1, wq + 4);
                gram_group_g4(c, rrp, st + 1024 + l_even, e_sw0, e_sw1, wq + 8);
                gram_group_g4(c, rrp, st + 1024 + l_odd, o_sw0, o_sw1, wq + 12);
            } else {
                if (meta.x > 0) gram_group_g4(c, rrp, st + l_even, e_sw0, e_sw1, wq);
                if (meta.x > 4) gram_group_g4(c, rrp, st + l_odd, o_sw0, o_sw1, wq + 4);
                if (meta.x > 8) gram_group_g4(c, rrp, st + 1024 + l_even, e_sw0, e_sw1, wq + 8);
                if (meta.x > 12) gram_group_g4(c, rrp, st + 1024 + l_odd, o_sw0, o_sw1, wq + 12);
            }
        }
        __syncwarp();                     // every lane is done reading slot h
        if (!meta.w) {                    // more stages of this item to come: refill the slot and go on
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
            continue;
        }
        // ---------------- tail: one item's Gram is complete; slot h's rows are its scratch until the refill at the end
        tail32_warp<0, false>(c, rrp, meta.y, st, sLF, srr0, p, lane, vecs);
        __syncwarp();                     // the scratch is free again
        issue_stage(h);
        h = (h + 1 == NS) ? 0 : h + 1;
    }
}

// =====================================================================================================================
// Heavy items (skew handling): an item with tens of thousands of ratings (ChEMBL's hottest target has 110 118) would be
// ONE warp's work in the stream kernel and set the duration of the whole sweep. Such items are cut into chunks of
// HEAVY_CHUNK ratings: heavy_gram32_kernel computes one partial Gram per chunk (a warp per chunk, DMMA, fragments
// straight from global memory), heavy_tail32_kernel adds an item's partials in chunk order (so the result does not
// depend on scheduling) and runs the same tail as the stream kernel. The stream kernel (SKIP instantiation) passes over
// the heavy items of its range.
// =====================================================================================================================
constexpr int HEAVY_CHUNK = 1024;
constexpr int HEAVY_PART = 24 * 32;           // doubles per partial: c[10][2] + rrp[4] per lane, in the DMMA layout

__global__ void __launch_bounds__(128) heavy_gram32_kernel(StreamArgs p, int nchunks, const int64_t *__restrict__ ch_p0,
                                                           const int64_t *__restrict__ ch_p1, double *__restrict__ partials)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ch = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ch >= nchunks) return;
    const int64_t p0 = ch_p0[ch], p1 = ch_p1[ch];
    double c[10][2], rrp[4];
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;
    for (int64_t q = p0; q < p1; q += 4) {
        const bool live = q + t < p1;
        double f[4], w = 0.0;
        if (live) {
            const double *row = p.other + (size_t)__ldg(p.rowidx + q + t) * 32 + g;
#pragma unroll
            for (int a = 0; a < 4; ++a) f[a] = __ldg(row + 8 * a);
            w = (__ldg(p.val + q + t) - p.mean_rating) * p.alpha;
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) f[a] = 0.0;
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], f[I], f[J]);
#pragma unroll
        for (int a = 0; a < 4; ++a) rrp[a] = fma(f[a], w, rrp[a]);
    }
    double *out = partials + (size_t)ch * HEAVY_PART + lane;
#pragma unroll
    for (int b = 0; b < 10; ++b) { out[(2 * b) * 32] = c[b][0]; out[(2 * b + 1) * 32] = c[b][1]; }
#pragma unroll
    for (int a = 0; a < 4; ++a) out[(20 + a) * 32] = rrp[a];
}

// one warp (one CTA) per heavy item; PROP: per-item prior precision (the tail reads it from p.propLambda, srr0 holds hp.mu)
template <bool PROP>
__global__ void __launch_bounds__(32) heavy_tail32_kernel(StreamArgs p, const int *__restrict__ hv_item, const int *__restrict__ hv_first,
                                                          const double *__restrict__ partials)
{
    __shared__ __align__(16) double sLF[32 * LFS + 32];
    __shared__ __align__(16) unsigned char scratch[V3_B_OFF + 256];
    const int lane = threadIdx.x;
    double *srr0 = sLF + 32 * LFS;
    for (int e = lane; e < 1024; e += 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    __syncwarp();
    {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + lane] * p.mu[j];
        srr0[lane] = PROP ? p.mu[lane] : s;
    }
    __syncwarp();
    const int h = blockIdx.x;
    double c[10][2], rrp[4];
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;
    for (int ch = hv_first[h]; ch < hv_first[h + 1]; ++ch) {       // fixed order: independent of scheduling
        const double *in = partials + (size_t)ch * HEAVY_PART + lane;
#pragma unroll
        for (int b = 0; b < 10; ++b) { c[b][0] += in[(2 * b) * 32]; c[b][1] += in[(2 * b + 1) * 32]; }
#pragma unroll
        for (int a = 0; a < 4; ++a) rrp[a] += in[(20 + a) * 32];
    }
    tail32_warp<0, PROP, TV_DEFAULT & 4>(c, rrp, hv_item[h], scratch, sLF, srr0, p, lane);
}

#ifdef BPMF_STREAM_PROBES               // experiments that were measured slower: not part of the product build
#include "../../bench_micro/stream_experiments.cuh"
#include "../../bench_micro/stream_roles.cuh"
#endif

// w[i] = (val[i] - mean_rating) * alpha: the same two roundings as in the kernels (a subtraction feeding a product cannot fuse)
__global__ void weights_kernel(const double *__restrict__ val, double *__restrict__ w, long long n, double mean_rating, double alpha)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        w[i] = __dmul_rn(__dsub_rn(val[i], mean_rating), alpha);
}

// The TV mask (chol3_block_column, issue_stage) of a kernel version. The product kernels (3, 12, 13, 16) run TV = 4 | 16:
// panel blocks of the factor from the trailing update's B fragments, stage indices held by the lanes that copy their rows
// (7.33 -> 7.19 ms, profiles/r02b_tune_tail_variants2.log). 24 = neither, the others are the measured alternatives.
constexpr int tail_variant(int VER)
{
    return VER == 15 ? (TV_DEFAULT | 1) : VER == 17 ? (TV_DEFAULT | 2) : VER == 19 ? (TV_DEFAULT | 8) : VER == 18 ? (16 | 128) : VER == 21 ? (4 | 128) : VER == 24 ? 128 :
           VER == 27 ? (4 | 16) : VER == 28 ? (TV_DEFAULT & ~256) : TV_DEFAULT;
}

template <int NS, int NW, int VER, int DBG = 0>
cudaError_t launch_cfg(bpmf_gpu_ctx *c, const StreamArgs &p, long long n)
{
    constexpr size_t smem = (size_t)NW * warp_bytes<NS>() + SHARED_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget");
#ifdef BPMF_STREAM_PROBES
    auto kern = VER >= 3 ? items_stream32v3_kernel<NS, NW, DBG, VER == 6, (VER == 7 ? 1 : VER == 8 ? 2 : VER == 9 ? 3 : 0), VER == 12 || VER == 16, VER == 13 || VER == 16, tail_variant(VER)> : items_stream32_kernel<NS, NW>;
#else
    static_assert(VER >= 3, "the v2 kernel is an experiment (stream_experiments.cuh)");
    auto kern = items_stream32v3_kernel<NS, NW, DBG, VER == 6, (VER == 7 ? 1 : VER == 8 ? 2 : VER == 9 ? 3 : 0), VER == 12 || VER == 16, VER == 13 || VER == 16, tail_variant(VER)>;
#endif
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = p.sms;                                    // persistent: one CTA per SM
    // small sweeps are claimed CLAIM_TAIL items at a time (bulk_end == from): enough CTAs for every claim to find a warp
    const long long need = (n + (long long)NW * CLAIM_TAIL - 1) / ((long long)NW * CLAIM_TAIL);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, NW * 32, smem, c->stream>>>(p);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// Sweep reductions for K = 32 (sample.cpp:359-362,379-381): sum of x x^T on the fp64 tensor cores, sum of x, and
// norm = trace. Fixed decomposition (STATS_BLOCKS x 8 warps, contiguous item chunks) and fixed summation order, so
// the result does not depend on how items were scheduled in the sweep, nor on the number of GPUs.
// Writes the same per-block partial layout as stats_partial_kernel (exact_kernels.cu): prod[K*K] | sum[K] | norm.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SW = 8;   // warps per stats block

// One block of the fixed STATS_BLOCKS decomposition, computed by a group of SW warps (tid = 0 .. SW * 32 - 1 inside the group,
// `bar` = the group's named barrier, sp = its SW x 672 doubles of shared memory); the partial goes to `partials` and to every
// peer's copy. The order of every sum is fixed by the block id alone.
__device__ __forceinline__ void group_sync(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(SW * 32) : "memory"); }

__device__ __forceinline__ void stats_block32(const double *__restrict__ items, int N, double *__restrict__ partials, int blk_id, int npeers,
                                              double *const *__restrict__ peers, double (*sp)[10 * 64 + 32], int tid, int bar)
{
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long nw = (long long)STATS_BLOCKS * SW, w = (long long)blk_id * SW + warp;
    long long chunk = (N + nw - 1) / nw;
    chunk = (chunk + 3) & ~3ll;
    const long long i0 = min((long long)N, w * chunk), i1 = min((long long)N, i0 + chunk);
    double c[10][2], sx[4];
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) sx[a] = 0.0;
    for (long long i = i0; i < i1; i += 4) {
        double f[4];
        const bool valid = i + t < i1;
        const double *row = items + (size_t)(i + t) * 32 + g;
#pragma unroll
        for (int a = 0; a < 4; ++a) f[a] = valid ? __ldg(row + 8 * a) : 0.0;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], f[I], f[J]);
#pragma unroll
        for (int a = 0; a < 4; ++a) sx[a] += f[a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        sx[a] += __shfl_xor_sync(FULL, sx[a], 1);
        sx[a] += __shfl_xor_sync(FULL, sx[a], 2);
    }
#pragma unroll
    for (int b = 0; b < 10; ++b) {
        sp[warp][b * 64 + g * 8 + 2 * t] = c[b][0];        // block b, element (g, 2t + e)
        sp[warp][b * 64 + g * 8 + 2 * t + 1] = c[b][1];
    }
    if (t == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) sp[warp][640 + 8 * a + g] = sx[a];
    }
    group_sync(bar);
    // the block's partial, first in registers (every sp[][] value is still needed), then in sp itself
    double acc4[1024 / (SW * 32)], accs = 0.0;
#pragma unroll
    for (int q = 0; q < 1024 / (SW * 32); ++q) {
        const int e = tid + q * SW * 32;
        int i = e & 31, k = e >> 5;                          // prod(i,k), column-major
        if (i < k) { const int x = i; i = k; k = x; }        // mirror of the lower triangle
        const int I = i >> 3, J = k >> 3;
        const int off = blk(I, J) * 64 + (i & 7) * 8 + (k & 7);
        double acc = 0.0;
#pragma unroll
        for (int ww = 0; ww < SW; ++ww) acc += sp[ww][off];
        acc4[q] = acc;
    }
    if (tid < 32) {
#pragma unroll
        for (int ww = 0; ww < SW; ++ww) accs += sp[ww][640 + tid];
    }
    group_sync(bar);
    double *so = &sp[0][0];                                  // 1024 + 32 + 1 doubles
#pragma unroll
    for (int q = 0; q < 1024 / (SW * 32); ++q) so[tid + q * SW * 32] = acc4[q];
    if (tid < 32) so[1024 + tid] = accs;
    group_sync(bar);
    if (tid == 0) {   // norm = sum of squared norms = trace of the outer-product sum
        double nn = 0.0;
        for (int d = 0; d < 32; ++d) nn += so[d * 33];
        so[1024 + 32] = nn;
    }
    group_sync(bar);
    const size_t base = (size_t)blk_id * (1024 + 32 + 1);
    for (int e = tid; e < 1024 + 32 + 1; e += SW * 32) {
        const double v = so[e];
        partials[base + e] = v;
        for (int q = 0; q < npeers; ++q) {
            double *dst = peers[q];
            if (dst && dst != partials) dst[base + e] = v;
        }
    }
    group_sync(bar);                      // sp is free for the group's next block
}

// Block b0 + blockIdx.x, one block per CTA: the reductions on the main stream (all SMs)
__global__ void __launch_bounds__(SW * 32) stats_partial32_kernel(const double *__restrict__ items, int N, double *__restrict__ partials, int b0,
                                                                  int npeers, double *const *__restrict__ peers)
{
    __shared__ double sp[SW][10 * 64 + 32];
    stats_block32(items, N, partials, b0 + (int)blockIdx.x, npeers, peers, sp, threadIdx.x, 0);
}

// The same blocks [b0, b0 + nb) by a FEW persistent CTAs of four groups each (grid = the SMs the item kernels leave free): the
// reductions on the auxiliary stream, under the other side's sweep. A wide grid would spread over every SM in the gap between
// two item kernels and keep the next one's CTAs (which need a whole SM each) waiting.
constexpr int SG = 4;   // groups per CTA
__global__ void __launch_bounds__(SG * SW * 32, 1) stats_partial32_persistent_kernel(const double *__restrict__ items, int N, double *__restrict__ partials,
                                                                                   int b0, int nb, int npeers, double *const *__restrict__ peers)
{
    extern __shared__ __align__(16) unsigned char stats_smem[];
    const int gi = threadIdx.x / (SW * 32), tid = threadIdx.x % (SW * 32);
    double (*sp)[10 * 64 + 32] = reinterpret_cast<double (*)[10 * 64 + 32]>(stats_smem) + (size_t)gi * SW;
    for (int vb = (int)blockIdx.x * SG + gi; vb < nb; vb += (int)gridDim.x * SG)
        stats_block32(items, N, partials, b0 + vb, npeers, peers, sp, tid, 1 + gi);
}

}  // namespace

cudaError_t launch_stats_partial32(bpmf_gpu_ctx *c, int side, int b0, int nb, cudaStream_t stream)
{
    SideDev &s = c->side[side];
    if (nb < 1) return cudaSuccess;
    if (stream != c->stream && c->stats_aux) {
        // auxiliary stream: as many CTAs as SMs are left free by the item kernels
        constexpr size_t smem = sizeof(double) * SG * SW * (10 * 64 + 32);
        const cudaError_t e = cudaFuncSetAttribute(stats_partial32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // (per device)
        if (e != cudaSuccess) return e;
        const int grid = std::max(1, std::min(c->reserve_sms, (nb + SG - 1) / SG));
        stats_partial32_persistent_kernel<<<grid, SG * SW * 32, smem, stream>>>(s.items, s.num, s.partials, b0, nb, s.n_stat_peers, s.stat_peers_dev);
    } else {
        stats_partial32_kernel<<<nb, SW * 32, 0, stream>>>(s.items, s.num, s.partials, b0, s.n_stat_peers, s.stat_peers_dev);
    }
    c->launches++;
    return cudaGetLastError();
}

// items per statistics block for K == 32: SW warps x a warp's chunk (a multiple of four)
int stats32_block_items(int num)
{
    const long long nw = (long long)STATS_BLOCKS * SW;
    long long chunk = (num + nw - 1) / nw;
    chunk = (chunk + 3) & ~3ll;
    return (int)(chunk * SW);
}

// 2-D tensor map of a latent matrix for the gather4 kernel: [num rows][32 doubles], box = 16 doubles x 1 row, 128-byte swizzle
static cudaError_t make_gather_map(const double *items, int num, CUtensorMap *out)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {32, (cuuint64_t)(num > 0 ? num : 1)};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {16, 1}, estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(items), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int NS, int NW, bool SKIP>
static cudaError_t launch_g4(bpmf_gpu_ctx *c, StreamArgs p, long long n, int num_other)
{
    constexpr size_t smem = (size_t)G4_SHARED + (size_t)NW * NS * G4_ROWS_BYTES + (size_t)NW * g4_aux_warp<NS>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    CUtensorMap tmap;
    cudaError_t e = make_gather_map(p.other, num_other, &tmap);
    if (e != cudaSuccess) return e;
    p.oob_row = num_other > 0 ? num_other : 1;
    auto kern = items_stream32g4_kernel<NS, NW, SKIP>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = p.sms;
    const long long need = (n + (long long)NW * CLAIM_TAIL - 1) / ((long long)NW * CLAIM_TAIL);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, NW * 32, smem, c->stream>>>(p, tmap);
    return cudaGetLastError();
}

static cudaError_t launch_stream_range(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha, int from, int to, bool skip_heavy)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    StreamArgs p;
    p.from = from; p.to = to; p.iter = iter; p.alpha = alpha; p.mean_rating = s.mean_rating;
    p.colptr = s.colptr; p.rowidx = s.rowidx; p.val = s.val;
    p.other = o.items; p.items = s.items;
    p.npeers = s.npeers; p.peers = s.peers_dev;
    p.mu = s.hp.mu; p.LambdaF = s.hp.LambdaF;
    p.work_counter = s.work_counter; p.err = c->d_err; p.zero_row = c->d_zero_row; p.heavy_thr = s.heavy_thr; p.propLambda = s.propLambda;
    p.oob_row = o.num;
    p.sms = item_sms(c, side);
    p.wval = nullptr;
    if (TV_DEFAULT & 128) {
        cudaError_t e0 = cudaSuccess;
        if (!s.wval) {
            e0 = cudaMalloc(&s.wval, sizeof(double) * (size_t)(s.nnz + 32));
            if (e0 != cudaSuccess) return e0;
            s.wval_valid = false;
        }
        if (!s.wval_valid || s.wval_alpha != alpha) {
            if ((e0 = cudaMemsetAsync(s.wval, 0, sizeof(double) * (size_t)(s.nnz + 32), c->stream)) != cudaSuccess) return e0;
            if (s.nnz > 0) weights_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(s.val, s.wval, (long long)s.nnz, s.mean_rating, alpha);
            s.wval_valid = true; s.wval_alpha = alpha;
        }
        p.wval = s.wval;
    }
    cudaError_t e = cudaMemsetAsync(s.work_counter, 0, 2 * sizeof(unsigned int), c->stream);
    if (e != cudaSuccess) return e;
    const long long n = (long long)to - from;
    if (n < 1) return cudaSuccess;
    // guided self-scheduling by default: a claim = remaining items / resident warps (4 quarters), between CLAIM_TAIL and CLAIM
    // (7.70 vs 7.71 ms on the whole of Synthetic A, 1.002 vs 1.050 ms on an eighth of it: profiles/r02_tune_gather4_guided*.log);
    // tuning value 9 = the fixed bulk / tail split of round 1
    p.guided = 0;
    {
        const int gq = c->stream_guided ? c->stream_guided : 4;
        if (gq != 9) p.guided = (int)std::min<long long>(0x7fffffff, std::max<long long>(1, (long long)gq * p.sms * 20 / 4));
    }
    {   // the last ~4 items per resident warp are handed out CLAIM_TAIL at a time (v3 kernel); bulk region is a multiple of CLAIM
        static const int tail_per_warp = [] { const char *v = getenv("BPMF_STREAM_TAIL"); return v ? atoi(v) : 8; }();
        long long tail_items = (long long)c->sm_count * 16 * tail_per_warp;
        long long bulk = n - tail_items;
        if (bulk < 0) bulk = 0;
        bulk -= bulk % CLAIM;
        p.bulk_end = from + (int)bulk;
        if (c->stream_tail >= 0) {
            tail_items = (long long)c->sm_count * 16 * c->stream_tail;
            bulk = n - tail_items; if (bulk < 0) bulk = 0; bulk -= bulk % CLAIM;
            p.bulk_end = from + (int)bulk;
        }
    }
    // tuning knob (bench_micro/tune_stream.py): "<version><stages><warps>", e.g. 3216 = v3, 2 stages x 16 warps.
    // Set with bpmf_gpu_debug_set_tuning or the environment variable BPMF_STREAM_CFG.
    static const int env_cfg = [] { const char *v = getenv("BPMF_STREAM_CFG"); return v ? atoi(v) : 0; }();
    const int cfg = c->stream_cfg ? c->stream_cfg : env_cfg;
    if (s.propLambda) {                   // per-item prior precisions (propagated posterior): v3, 2 stages x 16 warps
        e = skip_heavy ? launch_cfg<2, 16, 16>(c, p, n) : launch_cfg<2, 16, 13>(c, p, n);
        c->launches++;
        return e;
    }
    if (skip_heavy) {                     // v3, 2 stages x 20 warps, passing over the heavy items
        e = launch_cfg<2, 20, 12>(c, p, n);
        c->launches++;
        return e;
    }
#ifdef BPMF_STREAM_PROBES
    if (c->v5_gram_mask) {                // warp-role kernel (bpmf_gpu_debug_set_roles; bench_micro/stream_roles.cuh)
        V5Cfg v5{c->v5_gram_mask, c->v5_ns, c->v5_nslot};
        e = c->v5_nw == 24 ? launch_v5<24>(c, p, n, v5) : c->v5_nw == 16 ? launch_v5<16>(c, p, n, v5) : launch_v5<20>(c, p, n, v5);
        c->launches++;
        return e;
    }
#else
    if (c->v5_gram_mask) return cudaErrorNotSupported;      // an experiment: build with BPMF_STREAM_PROBES=1
#endif
    switch (cfg) {
    // 14<NS><NW>: TMA gather4 (cp.async.bulk.tensor.2d tile::gather4, UTMALDG) instead of cp.async
    case 15220: e = launch_cfg<2, 20, 15>(c, p, n); break;       // rank-one DMMA column steps in the LDL^T (see chol3_block_column)
    case 17220: e = launch_cfg<2, 20, 17>(c, p, n); break;       // + pivots and reciprocals through shared memory (7.38 vs 7.33 ms)
    case 19220: e = launch_cfg<2, 20, 19>(c, p, n); break;       // + rank-one DMMA column steps for the block columns 2 and 3 (7.355 vs 7.338 ms)
    case 18220: e = launch_cfg<2, 20, 18>(c, p, n); break;       // without the factor's panel blocks from the trailing update's B fragments
    case 21220: e = launch_cfg<2, 20, 21>(c, p, n); break;       // without the permuted ownership of the stage's indices
    case 24220: e = launch_cfg<2, 20, 24>(c, p, n); break;       // without either (7.33 ms)
    case 27220: e = launch_cfg<2, 20, 27>(c, p, n); break;       // with the ratings' weights computed per staged rating in every sweep
    case 28220: e = launch_cfg<2, 20, 28>(c, p, n); break;       // with the right-hand side's quad sums by shuffles instead of through shared memory
    case 14220: e = launch_g4<2, 20, false>(c, p, n, o.num); break;
    case 14216: e = launch_g4<2, 16, false>(c, p, n, o.num); break;
    case 14316: e = launch_g4<3, 16, false>(c, p, n, o.num); break;
    // 6<NS><NW>: v3 with the TMA bulk-copy gather (cp.async.bulk + mbarrier; measured slower, profiles/r01_tune_bulk_tma.log)
    case 6216: e = launch_cfg<2, 16, 6>(c, p, n); break;
    case 6220: e = launch_cfg<2, 20, 6>(c, p, n); break;
#ifdef BPMF_STREAM_PROBES
#include "../../bench_micro/stream_experiment_cases.inc"
#endif
    case 3216: e = launch_cfg<2, 16, 3>(c, p, n); break;
    default: e = launch_cfg<2, 20, 3>(c, p, n); break;   // v3, 2 stages x 20 warps
    }
    c->launches++;
    return e;
}

cudaError_t launch_items_stream32(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    // heavy items of [from, to): the stream kernel passes over them
    int first = 0, last = 0;
    while (first < s.n_heavy && s.h_heavy_item[first] < s.from) ++first;
    last = first;
    while (last < s.n_heavy && s.h_heavy_item[last] < s.to) ++last;
    if (first == last) return launch_stream_range(c, side, iter, alpha, s.from, s.to, false);
    {
        const cudaError_t e = launch_stream_range(c, side, iter, alpha, s.from, s.to, true);
        if (e != cudaSuccess) return e;
    }
    StreamArgs p;
    p.from = s.from; p.to = s.to; p.iter = iter; p.alpha = alpha; p.mean_rating = s.mean_rating;
    p.colptr = s.colptr; p.rowidx = s.rowidx; p.val = s.val;
    p.other = o.items; p.items = s.items;
    p.npeers = s.npeers; p.peers = s.peers_dev;
    p.mu = s.hp.mu; p.LambdaF = s.hp.LambdaF;
    p.work_counter = s.work_counter; p.err = c->d_err; p.zero_row = c->d_zero_row; p.bulk_end = s.to; p.guided = 0; p.heavy_thr = s.heavy_thr;
    p.propLambda = s.propLambda;
    const int ch0 = s.h_heavy_first[first], ch1 = s.h_heavy_first[last];
    heavy_gram32_kernel<<<(ch1 - ch0 + 3) / 4, 128, 0, c->stream>>>(p, ch1 - ch0, s.hv_p0 + ch0, s.hv_p1 + ch0,
                                                                     s.hv_partials + (size_t)ch0 * HEAVY_PART);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (s.propLambda) heavy_tail32_kernel<true><<<last - first, 32, 0, c->stream>>>(p, s.hv_item + first, s.hv_first + first, s.hv_partials);
    else heavy_tail32_kernel<false><<<last - first, 32, 0, c->stream>>>(p, s.hv_item + first, s.hv_first + first, s.hv_partials);
    c->launches++;
    return cudaGetLastError();
}

int heavy_chunk_size() { return HEAVY_CHUNK; }
int heavy_partial_doubles() { return HEAVY_PART; }

}  // namespace bpmf
