// stream_experiments.cuh — kernel variants that were built, verified bit-identical to the product kernel and MEASURED
// SLOWER (profiles/r01_tune_*.log). They are kept as evidence and as starting points, and are compiled only with
// BPMF_STREAM_PROBES=1 (python -m bpmf_b200.build); the product path never launches them.
//   v2  LL^T tail with per-column rsqrt, double-batch index prefetch               8.72 ms   (profiles/r01_tune_v3.log)
//   v4  Gram warps + tail warps (warp specialisation, shared-memory handoff)        8.22 ms   (r01_tune_v4_warp_specialised.log)
//   v7  MMA warps + producer warps + tail warps, one DMMA-issuing warp / scheduler  9.9 ms    (r01_tune_v7_three_roles_*.log)
//   v8  two items per warp in the tail (two interleaved dependency chains)          9.48 ms   (r01_tune_v8_two_items_per_warp.log)
// Included by stream_kernel.cu inside namespace bpmf::{anonymous}; uses its helpers.
#pragma once
// v2 tail scratch: L (packed, 528 doubles) in the consumed stage, or, before that, zy | zr | b
constexpr int LPACK = 528;                  // packed lower triangle, column-major
constexpr int ZY_OFF = 0, ZR_OFF = 256, B_OFF = 512;
constexpr __host__ __device__ int col_off(int k) { return 32 * k - ((k * (k - 1)) / 2); }
static_assert(LPACK * 8 <= STAGE_BYTES && B_OFF + 256 <= STAGE_BYTES, "tail scratch must fit in one stage");
__device__ __forceinline__ void mbar_arrive(uint32_t mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar) : "memory");
}
// the mbarrier gets one arrival from this thread once all cp.async it has issued so far have landed (count pre-accounted)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t mbar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(mbar) : "memory");
}
// One block column KB of the right-looking blocked Cholesky on the accumulator layout (bench_micro/emulate_block_chol.py
// is the lane-level model of this function).
template <int KB>
__device__ __forceinline__ void chol_block_column(double (&c)[10][2], double &myrs, bool &ok, int lane, int t)
{
    constexpr int D = blk(KB, KB);
#pragma unroll 1
    for (int k2 = 0; k2 < 4; ++k2) {
        const bool own = (t == k2);
        const int qsrc = (lane & ~3) | k2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = 2 * k2 + e;
            const double p = __shfl_sync(FULL, c[D][e], 4 * k + k2);       // pivot: lane (g = k, t = k2), register e
            if (!(p > 0.0)) ok = false;                                     // Eigen LLT: pivot <= 0 -> "Cholesky failed"
            const double rs = rsqrt(p);
            if (lane == 8 * KB + k) myrs = rs;
            const double sc = own ? rs : 1.0;
#pragma unroll
            for (int I = KB; I < 4; ++I) c[blk(I, KB)][e] *= sc;            // column k of L (diagonal entry becomes sqrt(p))
            // L[2t][k], L[2t+1][k]; zeroed where this lane's column is not right of k, so the updates need no predicate
            double bl0 = __shfl_sync(FULL, c[D][e], 4 * (2 * t) + k2);
            double bl1 = __shfl_sync(FULL, c[D][e], 4 * (2 * t + 1) + k2);
            bl0 = (2 * t > k) ? -bl0 : 0.0;
            bl1 = (2 * t + 1 > k) ? -bl1 : 0.0;
#pragma unroll
            for (int I = KB; I < 4; ++I) {
                const double a = __shfl_sync(FULL, c[blk(I, KB)][e], qsrc);   // L[8I+g][k]
                c[blk(I, KB)][0] = fma(a, bl0, c[blk(I, KB)][0]);
                c[blk(I, KB)][1] = fma(a, bl1, c[blk(I, KB)][1]);
            }
        }
    }
    // trailing update A(I,J) -= L(I,KB) L(J,KB)^T for KB < J <= I on the tensor cores
    if (KB < 3) {
        double fr[4][2];
#pragma unroll
        for (int I = KB + 1; I < 4; ++I)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int src = (lane & ~3) | (2 * kk + (t >> 1));
                const double v0 = __shfl_sync(FULL, c[blk(I, KB)][0], src);
                const double v1 = __shfl_sync(FULL, c[blk(I, KB)][1], src);
                fr[I][kk] = (t & 1) ? v1 : v0;                              // L(I,KB)[g][4kk + t]
            }
#pragma unroll
        for (int I = KB + 1; I < 4; ++I)
#pragma unroll
            for (int J = KB + 1; J <= I; ++J)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) dmma884(c[blk(I, J)][0], c[blk(I, J)][1], -fr[I][kk], fr[J][kk]);
    }
}

template <int NS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) items_stream32_kernel(StreamArgs p)
{
    constexpr int WARP_BYTES = warp_bytes<NS>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    unsigned char *wbase = smem_raw + SHARED_BYTES + (size_t)warp * WARP_BYTES;
    const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);

    for (int e = tid; e < 1024; e += NW * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    // ---------------- fetch-side state (warp-uniform unless noted); positions are relative to the group's first rating
    int g_base = 0, g_n = 0, f_it = 0;
    int cpr = 0;                          // per lane: colptr[g_base + lane] - colptr[g_base]
    int f_pos = 0, f_end = 0, f_start = 0, g_end = 0, b_base = 0;
    const int32_t *g_idx = p.rowidx;      // rowidx / val at the group's first rating
    const double *g_val = p.val;
    int32_t b_idx = 0, nb_idx = 0;        // per lane: index of stream position b_base + lane / b_base + 32 + lane
    double b_w = 0.0, nb_w = 0.0;         // per lane: the rating values of the same positions (used long after the load)
    bool f_done = false;
    const unsigned char *src_lane = reinterpret_cast<const unsigned char *>(p.other) + (lane & 15) * 16;   // this lane's 16 B of a row
    const uint32_t dst_lane = wbase_s + (lane >> 4) * ROWB + (lane & 15) * 16;

    auto load_batch = [&](int base, int32_t &idx, double &w) {
        const int q = base + lane;
        idx = 0; w = 0.0;
        if (q < g_end) {
            idx = __ldg(g_idx + q);
            w = __ldg(g_val + q);         // raw: any arithmetic here would wait for the load right away
        }
    };
    auto claim = [&]() {
        int base = 0;
        if (lane == 0) base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.to) { f_done = true; return; }
        g_base = base;
        g_n = min(CLAIM, p.to - base);
        const int64_t c0 = __ldg(p.colptr + base);
        cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
        g_idx = p.rowidx + c0;
        g_val = p.val + c0;
        f_it = 0;
        f_start = f_pos = 0;
        f_end = __shfl_sync(FULL, cpr, 1);
        g_end = __shfl_sync(FULL, cpr, g_n);
        b_base = 0;
        load_batch(0, b_idx, b_w);
        load_batch(32, nb_idx, nb_w);
    };
    // fill ring slot `slot` with the next (at most SR) ratings of the current item; exactly one commit_group per call
    auto issue_stage = [&](int slot) {
        const uint32_t st = dst_lane + slot * STAGE_BYTES;
        unsigned char *stg = wbase + slot * STAGE_BYTES;
        if (f_done) {
            if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(-1, 0, 0, 0);
            cp_async_commit();
            return;
        }
        const int off = f_pos - b_base;                  // 0..31: the stage may run into the prefetched batch
        const int n = min(SR, f_end - f_pos);
#pragma unroll
        for (int i = 0; i < SR / 2; ++i) {
            const int q = off + 2 * i + (lane >> 4);
            const int j0 = __shfl_sync(FULL, b_idx, q & 31);
            const int j1 = __shfl_sync(FULL, nb_idx, q & 31);
            const unsigned j = (unsigned)((q < 32) ? j0 : j1);
            // rows past the item's end are zero-filled (src-size 0 reads nothing; j is still a valid row)
            cp_async16(st + 2 * i * ROWB, src_lane + (size_t)j * 256, (2 * i + (lane >> 4) < n) ? 16 : 0);
        }
        {
            const int q = off + lane;
            const double w0 = __shfl_sync(FULL, b_w, q & 31);
            const double w1 = __shfl_sync(FULL, nb_w, q & 31);
            // rr weight (v - mean_rating) * alpha (sample.cpp:255)
            if (lane < SR) reinterpret_cast<double *>(stg + W_OFF)[lane] = (lane < n) ? (((q < 32) ? w0 : w1) - p.mean_rating) * p.alpha : 0.0;
        }
        const int last = (f_pos + n == f_end);
        if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(n, g_base + f_it, f_pos == f_start, last);
        cp_async_commit();
        f_pos += n;
        if (f_pos - b_base >= 32) {       // rotate to the prefetched batch and start fetching the one after it
            b_base += 32;
            b_idx = nb_idx; b_w = nb_w;
            load_batch(b_base + 32, nb_idx, nb_w);
        }
        if (last) {
            ++f_it;
            if (f_it >= g_n) claim();
            else { f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1); }
        }
    };

    claim();
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
        cp_async_wait<NS - 1>();
        __syncwarp();
        unsigned char *stg = wbase + h * STAGE_BYTES;
        const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
        if (meta.x < 0) break;
        // ---------------- Gram + rhs of this stage (computeMuLambda, sample.cpp:251-257) ----------------
        {
            const unsigned char *row = stg + t * ROWB + g * 8;
            const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
            if (meta.x > 0) gram_group(c, rrp, row, wq);
            if (meta.x > 4) gram_group(c, rrp, row + 4 * ROWB, wq + 4);
            if (meta.x > 8) gram_group(c, rrp, row + 8 * ROWB, wq + 8);
            if (meta.x > 12) gram_group(c, rrp, row + 12 * ROWB, wq + 12);
        }
        __syncwarp();                     // every lane is done reading slot h
        if (!meta.w) {                    // more stages of this item to come: refill the slot and go on
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
            continue;
        }
        // ---------------- tail: one item's Gram is complete; slot h is its scratch until the refill at the end ---------
        const int idx = meta.y;
        double *zy = reinterpret_cast<double *>(stg + ZY_OFF), *zr = reinterpret_cast<double *>(stg + ZR_OFF);
        double *wb = reinterpret_cast<double *>(stg + B_OFF), *Lp = reinterpret_cast<double *>(stg);
        // the K normals of this item: rng_set_pos((idx+1)*K*(iter+1)) (sample.cpp:266). Accepted polar attempts are numbered
        // by ballot; lane n then finishes normal n (one log / sqrt / divide per lane instead of one per attempt).
        {
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
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 1);
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) wb[8 * a + g] = srr0[8 * a + g] + rrp[a];
        }
        __syncwarp();
        const double z = __dmul_rn(zy[lane], polar_mult(zr[lane]));
        double bb = wb[lane];
        // MM = LambdaF + alpha * G (sample.cpp:297-298), in place in the accumulator layout
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    c[blk(I, J)][e] = fma(p.alpha, c[blk(I, J)][e], sLF[(8 * J + 2 * t + e) * LFS + 8 * I + g]);   // LambdaF(i,k), i >= k
        // chol.compute(MM) (sample.cpp:306)
        double myrs = 0.0;
        bool ok = true;
        chol_block_column<0>(c, myrs, ok, lane, t);
        chol_block_column<1>(c, myrs, ok, lane, t);
        chol_block_column<2>(c, myrs, ok, lane, t);
        chol_block_column<3>(c, myrs, ok, lane, t);
        __syncwarp();                     // zy / zr / wb have been read by everyone: L may overwrite them
        // L -> shared memory, packed by columns: element (i,k), i >= k, at col_off(k) + i - k
        {
            double *lq = Lp + g - 2 * t;                    // + col_off(k) + 8 (I - J) - e per element
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = 8 * J + 2 * t + e;
                        if (I > J || g >= 2 * t + e) lq[(32 * k - ((k * (k - 1)) >> 1)) + 8 * (I - J) - e] = c[blk(I, J)][e];
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
            // lane j owns row j; one broadcast per step. Lanes that are already solved are simply not updated.
            {
                const double *lf = Lp + lane;              // element (lane, k) at lf[col_off(k) - k]
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const double yk = __shfl_sync(FULL, bb * myrs, k);
                    if (lane > k) bb = fma(-lf[col_off(k) - k], yk, bb);
                }
            }
            double yv = fma(bb, myrs, z);                  // y + z
            {
                const double *lb = Lp + (32 * lane - ((lane * (lane - 1)) >> 1)) - lane;   // element (i, lane) at lb[i]
#pragma unroll
                for (int i = 31; i >= 0; --i) {
                    const double xi = __shfl_sync(FULL, yv * myrs, i);
                    if (lane < i) yv = fma(-lb[i], xi, yv);
                }
            }
            const double xv = yv * myrs;
            // items().col(idx) = rr (sample.cpp:324); push to the peer replicas (replaces send_item, :370)
            p.items[(size_t)idx * 32 + lane] = xv;
            for (int pr = 0; pr < p.npeers; ++pr) {
                double *dst = p.peers[pr];
                if (dst && dst != p.items) dst[(size_t)idx * 32 + lane] = xv;
            }
        } else if (lane == 0) {           // THROWERROR("Cholesky failed") (sample.cpp:308): reported through the error word
            atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
        }
        __syncwarp();                     // the scratch is free again
        issue_stage(h);
        h = (h + 1 == NS) ? 0 : h + 1;
    }
    cp_async_wait<0>();
}

// =====================================================================================================================
// Version 4: the same arithmetic as v3 with the two phases in SPECIALISED WARPS of one persistent CTA.
//   Gram warps (NG)  gather ring + DMMA Gram only. A finished item (the 20 accumulator doubles + 4 rhs partials per lane,
//                    6 KB, still in the DMMA layout) is handed to a free tail slot in shared memory.
//   tail warps (NT)  each owns one slot: wait for FULL, pull the accumulators into registers, release the slot, then
//                    normals + LDL^T + solves + store exactly as v3, in a private scratch.
// Why: DMMA and scalar fp64 share one pipe per scheduler and a DMMA holds it for 16 cycles, so in v3 every dependent
// scalar fp64 instruction of a tail queues behind another warp's Gram burst (math_pipe_throttle is the top stall) and the
// two phases add up instead of overlapping (probes: Gram only 4.74 ms, tail only 5.26 ms, both 8.06 ms). With separate
// warps the tails' few latency-critical instructions and the Grams' bulk DMMAs interleave; TAIL_HIGH puts the tail
// warps at the high warp ids, which the issue arbiter favours.
// Slot protocol (state word per slot in shared memory): 0 EMPTY -> 1 RESERVED (a Gram warp won the CAS and is writing)
// -> 2 FULL -> 0 ...; 3 EXIT is set by the last Gram warp to finish, only on EMPTY slots.
// =====================================================================================================================
constexpr int SLOT_ROWS = 24;                                  // c[10][2] + rrp[4]
constexpr int SLOT_BYTES = SLOT_ROWS * 32 * 8 + 16;            // + item index
constexpr int TSCRATCH_BYTES = V3_B_OFF + 256;                 // Lu (packed) | zy | zr | b
template <int NG, int NT, int NS>
constexpr __host__ __device__ size_t v4_smem_bytes()
{
    return (size_t)SHARED_BYTES + 64 + (size_t)NG * NS * STAGE_BYTES + (size_t)NT * (SLOT_BYTES + TSCRATCH_BYTES);
}

template <int NG, int NT, int NS, bool TAIL_HIGH, int DBG>
__global__ void __launch_bounds__((NG + NT) * 32, 1) items_stream32v4_kernel(StreamArgs p)
{
    static_assert(NT + 1 <= 16, "state words live in 64 bytes");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    volatile int *sstate = reinterpret_cast<volatile int *>(smem_raw + SHARED_BYTES);   // [NT] slots, [NT] = finished Gram warps
    unsigned char *rings = smem_raw + SHARED_BYTES + 64;
    unsigned char *slots = rings + (size_t)NG * NS * STAGE_BYTES;
    unsigned char *scratch = slots + (size_t)NT * SLOT_BYTES;

    for (int e = tid; e < 1024; e += (NG + NT) * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    if (tid < 16) sstate[tid] = 0;
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    const bool is_tail = TAIL_HIGH ? (warp >= NG) : (warp < NT);

    if (!is_tail) {
        // =============================================== Gram warp ===============================================
        const int gi = TAIL_HIGH ? warp : warp - NT;
        unsigned char *wbase = rings + (size_t)gi * NS * STAGE_BYTES;
        const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);
        int g_base = 0, g_n = 0, f_it = 0;
        int cpr = 0;                          // per lane: colptr[g_base + lane] - colptr[g_base]
        int f_pos = 0, f_end = 0, f_start = 0, g_end = 0;
        const int32_t *g_idx = p.rowidx;
        const double *g_val = p.val;
        int32_t n_idx = 0;                    // per lane: index / value of stream position f_pos + (lane & 15): the NEXT stage's
        double n_w = 0.0;
        bool f_done = false;
        const unsigned char *src_lane = reinterpret_cast<const unsigned char *>(p.other) + (lane & 15) * 16;
        const uint32_t dst_lane = wbase_s + (lane >> 4) * ROWB + (lane & 15) * 16;
        const int half = lane >> 4;
        int hint = (gi * NT) / NG;            // where this warp starts looking for an empty slot

        auto load_next = [&]() {
            const int q = f_pos + (lane & 15);
            n_idx = 0; n_w = 0.0;
            if (q < g_end) {
                n_idx = __ldg(g_idx + q);
                n_w = __ldg(g_val + q);
            }
        };
        auto claim = [&]() {
            int base = 0;
            if (lane == 0) base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
            base = __shfl_sync(FULL, base, 0);
            if (base >= p.to) { f_done = true; return; }
            g_base = base;
            g_n = min(CLAIM, p.to - base);
            const int64_t c0 = __ldg(p.colptr + base);
            cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
            g_idx = p.rowidx + c0;
            g_val = p.val + c0;
            f_it = 0;
            f_start = f_pos = 0;
            f_end = __shfl_sync(FULL, cpr, 1);
            g_end = __shfl_sync(FULL, cpr, g_n);
        };
        auto issue_stage = [&](int slot) {
            const uint32_t st = dst_lane + slot * STAGE_BYTES;
            unsigned char *stg = wbase + slot * STAGE_BYTES;
            if (f_done) {
                if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(-1, 0, 0, 0);
                cp_async_commit();
                return;
            }
            const int n = min(SR, f_end - f_pos);
            const int nn = n - half;
#pragma unroll
            for (int i = 0; i < SR / 2; ++i) {
                const unsigned j = (unsigned)__shfl_sync(FULL, n_idx, 2 * i + half);
                if (!(DBG & 4)) cp_async16(st + 2 * i * ROWB, src_lane + (size_t)j * 256, (2 * i < nn) ? 16 : 0);
            }
            if (lane < SR) reinterpret_cast<double *>(stg + W_OFF)[lane] = (lane < n) ? (n_w - p.mean_rating) * p.alpha : 0.0;
            const int last = (f_pos + n == f_end);
            if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(n, g_base + f_it, f_pos == f_start, last);
            cp_async_commit();
            f_pos += n;
            if (last) {
                ++f_it;
                if (f_it >= g_n) claim();
                else { f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1); }
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
            cp_async_wait<NS - 1>();
            __syncwarp();
            unsigned char *stg = wbase + h * STAGE_BYTES;
            const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
            if (meta.x < 0) break;
            {
                const unsigned char *row = stg + t * ROWB + g * 8;
                const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
                if (!(DBG & 2)) {
                    if (meta.x > 0) gram_group(c, rrp, row, wq);
                    if (meta.x > 4) gram_group(c, rrp, row + 4 * ROWB, wq + 4);
                    if (meta.x > 8) gram_group(c, rrp, row + 8 * ROWB, wq + 8);
                    if (meta.x > 12) gram_group(c, rrp, row + 12 * ROWB, wq + 12);
                } else if (meta.x > 0) {
                    c[0][0] += *reinterpret_cast<const double *>(row) * *wq;
                }
            }
            __syncwarp();                     // every lane is done reading slot h
            if (meta.w) {
                // ---- the item's Gram is complete: hand it to a tail warp
                int s = 0;
                if (lane == 0) {
                    s = hint;
                    for (;;) {
                        if (atomicCAS(const_cast<int *>(sstate) + s, 0, 1) == 0) break;
                        s = (s + 1 == NT) ? 0 : s + 1;
                        if (s == hint) __nanosleep(64);
                    }
                }
                s = __shfl_sync(FULL, s, 0);
                hint = (s + 1 == NT) ? 0 : s + 1;
                double *slot = reinterpret_cast<double *>(slots + (size_t)s * SLOT_BYTES) + lane;
#pragma unroll
                for (int b = 0; b < 10; ++b) {
                    slot[(2 * b) * 32] = c[b][0];
                    slot[(2 * b + 1) * 32] = c[b][1];
                    c[b][0] = 0.0; c[b][1] = 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) { slot[(20 + a) * 32] = rrp[a]; rrp[a] = 0.0; }
                if (lane == 0) *reinterpret_cast<int *>(slots + (size_t)s * SLOT_BYTES + SLOT_ROWS * 32 * 8) = meta.y;
                __syncwarp();
                if (lane == 0) { __threadfence_block(); sstate[s] = 2; }
            }
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
        }
        cp_async_wait<0>();
        if (lane == 0) {
            const int done = atomicAdd(const_cast<int *>(sstate) + NT, 1) + 1;
            if (done == NG)                   // the last Gram warp tells every tail warp to leave once its slot is empty
                for (int s = 0; s < NT; ++s)
                    while (atomicCAS(const_cast<int *>(sstate) + s, 0, 3) != 0) __nanosleep(64);
        }
        return;
    }

    // ================================================= tail warp =================================================
    const int ti = TAIL_HIGH ? warp - NG : warp;
    const double *slot = reinterpret_cast<const double *>(slots + (size_t)ti * SLOT_BYTES) + lane;
    unsigned char *stg = scratch + (size_t)ti * TSCRATCH_BYTES;
    double *zy = reinterpret_cast<double *>(stg + V3_ZY_OFF), *zr = reinterpret_cast<double *>(stg + V3_ZR_OFF);
    double *wb = reinterpret_cast<double *>(stg + V3_B_OFF), *Lp = reinterpret_cast<double *>(stg);
#pragma unroll 1
    for (;;) {
        int st = 0;
        if (lane == 0) {
            while ((st = sstate[ti]) < 2) __nanosleep(32);
            __threadfence_block();
        }
        st = __shfl_sync(FULL, st, 0);
        if (st == 3) break;
        double c[10][2];
        double rrp[4];
#pragma unroll
        for (int b = 0; b < 10; ++b) { c[b][0] = slot[(2 * b) * 32]; c[b][1] = slot[(2 * b + 1) * 32]; }
#pragma unroll
        for (int a = 0; a < 4; ++a) rrp[a] = slot[(20 + a) * 32];
        const int idx = *reinterpret_cast<const int *>(slots + (size_t)ti * SLOT_BYTES + SLOT_ROWS * 32 * 8);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sstate[ti] = 0; }   // the slot can take the next item while this one is solved
        if (DBG & 1) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 10; ++b) acc += c[b][0] + c[b][1];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc += rrp[a];
            p.items[(size_t)idx * 32 + lane] = acc;
            continue;
        }
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
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 1);
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) wb[8 * a + g] = srr0[8 * a + g] + rrp[a];
        }
        __syncwarp();
        const double z = (DBG & 8) ? 0.25 * lane : __dmul_rn(zy[lane], polar_mult(zr[lane]));
        double bb = wb[lane];
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    c[blk(I, J)][e] = fma(p.alpha, c[blk(I, J)][e], sLF[(8 * J + 2 * t + e) * LFS + 8 * I + g]);
        double myd = 1.0, myrinv = 1.0;
        bool ok = true;
        chol3_block_column<0>(c, myd, myrinv, lane, t);
        chol3_block_column<1>(c, myd, myrinv, lane, t);
        chol3_block_column<2>(c, myd, myrinv, lane, t);
        chol3_block_column<3>(c, myd, myrinv, lane, t);
        ok = __all_sync(FULL, myd > 0.0);
        const double myrs = rsqrt(myd);
        {
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * J + 2 * t + e;
                    const double rk = __shfl_sync(FULL, myrinv, k);
                    double *lq = Lp + (31 * k - ((k * (k - 1)) >> 1)) + g - k - 1;
#pragma unroll
                    for (int I = J; I < 4; ++I)
                        if (I > J || g > 2 * t + e) lq[8 * I] = c[blk(I, J)][e] * rk;
                }
        }
        __syncwarp();
        if (ok) {
            {
                const double *lf = Lp + lane - 1;
#pragma unroll
                for (int k = 0; k < 31; ++k) {
                    const double yk = __shfl_sync(FULL, bb, k);
                    if (lane > k) bb = fma(-lf[col_off1(k) - k], yk, bb);
                }
            }
            double yv = fma(bb, myrinv, myrs * z);
            {
                const double *lb = Lp + (31 * lane - ((lane * (lane - 1)) >> 1)) - lane - 1;
#pragma unroll
                for (int i = 31; i >= 1; --i) {
                    const double xi = __shfl_sync(FULL, yv, i);
                    if (lane < i) yv = fma(-lb[i], xi, yv);
                }
            }
            p.items[(size_t)idx * 32 + lane] = yv;
            for (int pr = 0; pr < p.npeers; ++pr) {
                double *dst = p.peers[pr];
                if (dst && dst != p.items) dst[(size_t)idx * 32 + lane] = yv;
            }
        } else if (lane == 0) {
            atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
        }
        __syncwarp();                         // the scratch is free again
    }
}

template <int NG, int NT, int NS, bool TAIL_HIGH, int DBG = 0>
cudaError_t launch_v4(bpmf_gpu_ctx *c, const StreamArgs &p, long long n)
{
    constexpr size_t smem = v4_smem_bytes<NG, NT, NS>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = items_stream32v4_kernel<NG, NT, NS, TAIL_HIGH, DBG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = c->sm_count;                              // persistent: one CTA per SM
    const long long need = (n + (long long)NG * CLAIM - 1) / ((long long)NG * CLAIM);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, (NG + NT) * 32, smem, c->stream>>>(p);
    return cudaGetLastError();
}

// =====================================================================================================================
// Version 7: three warp roles per scheduler, built on what bench_micro/fp64_latency.cu measured (profiles/
// r01_fp64_latency_microbench.txt): a dependent scalar-fp64 chain runs at 8 cycles/instruction alone, 14.6 with ONE warp
// of its scheduler streaming DMMAs and 56 with TWO. So exactly one warp per scheduler issues Gram DMMAs, and its
// instruction stream is kept (almost) pure DMMA so that it alone saturates the pipe:
//   MMA warps      (NM = 4, warp ids 0..3, one per scheduler)  poll a stage flag, fragment loads + DMMAs + rhs, free the
//                  stage, hand a finished item to a tail slot (as v4)
//   producer warps (NM, warp ids 4..7)  the fetch state machine of v3 for "their" MMA warp: index/value prefetch, cp.async
//                  gather into the MMA warp's ring, publish a stage when its copies have landed
//   tail warps     (NT, warp ids 8..)   exactly v4's tail warps
// Stage flag: 0 = empty (producer may fill), 1 = full (MMA warp may read).
// =====================================================================================================================
template <int NM, int NT, int NS>
constexpr __host__ __device__ size_t v7_smem_bytes()
{
    return (size_t)SHARED_BYTES + 64 + 128 + 512 + (size_t)NM * NS * STAGE_BYTES + (size_t)NT * (SLOT_BYTES + TSCRATCH_BYTES);
}

template <int NM, int NT, int NS, int DBG>
__global__ void __launch_bounds__((2 * NM + NT) * 32, 1) items_stream32v7_kernel(StreamArgs p)
{
    static_assert(NT + 1 <= 16 && NM * NS <= 28 && NM <= 4, "slot states live in 64 bytes, stage flags in 128");
    constexpr int NG = NM;                                          // for the tail-warp code shared with v4
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    volatile int *sstate = reinterpret_cast<volatile int *>(smem_raw + SHARED_BYTES);        // [NT] slots, [NT] = finished MMA warps
    volatile int *sflag = reinterpret_cast<volatile int *>(smem_raw + SHARED_BYTES + 64);    // [NM][NS] stage flags
    // mbarriers: [NM][NS] "stage empty" (MMA warp -> producer) at +0, [NT] "slot full" (MMA warp -> tail warp) at +256
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem_raw + SHARED_BYTES + 64 + 128);
    unsigned char *rings = smem_raw + SHARED_BYTES + 64 + 128 + 512;
    unsigned char *slots = rings + (size_t)NM * NS * STAGE_BYTES;
    unsigned char *scratch = slots + (size_t)NT * SLOT_BYTES;

    for (int e = tid; e < 1024; e += (2 * NM + NT) * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    if (tid < 16) sstate[tid] = 0;
    if (tid < 32) sflag[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NM * NS; ++i) mbar_init(bars + i * 8, 1);
        for (int i = 0; i < NT; ++i) mbar_init(bars + 256 + i * 8, 1);
    }
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    if (warp >= NM && warp < 2 * NM) {
        // ============================================= producer warp =============================================
        // Plans stages PF ahead of issuing them: the index / value loads of a stage are in flight for PF stage-issues
        // (the producer has no arithmetic to hide them behind).
        constexpr int PF = 3;
        const int gi = warp - NM;
        unsigned char *wbase = rings + (size_t)gi * NS * STAGE_BYTES;
        const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);
        int g_base = 0, g_n = 0, f_it = 0;
        int cpr = 0;
        int f_pos = 0, f_end = 0, f_start = 0, g_end = 0;
        const int32_t *g_idx = p.rowidx;
        const double *g_val = p.val;
        bool f_done = false;
        const unsigned char *src_lane = reinterpret_cast<const unsigned char *>(p.other) + (lane & 15) * 16;
        const uint32_t dst_lane = wbase_s + (lane >> 4) * ROWB + (lane & 15) * 16;
        const int half = lane >> 4;

        struct Plan { int n, item, first, last; int32_t idx; double w; };
        auto claim = [&]() {
            int base = 0;
            if (lane == 0) base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
            base = __shfl_sync(FULL, base, 0);
            if (base >= p.to) { f_done = true; return; }
            g_base = base;
            g_n = min(CLAIM, p.to - base);
            const int64_t c0 = __ldg(p.colptr + base);
            cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
            g_idx = p.rowidx + c0;
            g_val = p.val + c0;
            f_it = 0;
            f_start = f_pos = 0;
            f_end = __shfl_sync(FULL, cpr, 1);
            g_end = __shfl_sync(FULL, cpr, g_n);
        };
        // next stage of the stream: its extent, and the loads of its (at most SR) indices and values
        auto plan_next = [&]() {
            Plan q;
            q.idx = 0; q.w = 0.0;
            if (f_done) { q.n = -1; q.item = 0; q.first = 0; q.last = 0; return q; }
            q.n = min(SR, f_end - f_pos);
            q.item = g_base + f_it;
            q.first = (f_pos == f_start);
            q.last = (f_pos + q.n == f_end);
            const int e = f_pos + (lane & 15);
            if ((lane & 15) < q.n) {
                q.idx = __ldg(g_idx + e);
                q.w = __ldg(g_val + e);
            }
            f_pos += q.n;
            if (q.last) {
                ++f_it;
                if (f_it >= g_n) claim();
                else { f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1); }
            }
            return q;
        };

        if (lane == 0) {
#pragma unroll 1
            for (int s = 0; s < NS; ++s) mbar_init(wbase_s + s * STAGE_BYTES + MBAR_OFF, 32);
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sflag[28 + gi] = 1; }   // the MMA warp may now wait on the barriers
        claim();
        Plan plan[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) plan[k] = plan_next();
        int s_issue = 0, issued = 0;
        unsigned ephase = 0;                  // bit s: parity of the "empty" phase of stage s to wait for next
        bool done = false;
#pragma unroll 1
        while (!done) {
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                if (done) break;
                const Plan cur = plan[k];
                plan[k] = plan_next();
                // wait until the MMA warp has released the stage (the first NS uses find it empty)
                if (issued >= NS) {
                    mbar_wait(bars + (gi * NS + s_issue) * 8, (ephase >> s_issue) & 1u);
                    ephase ^= 1u << s_issue;
                }
                ++issued;
                __syncwarp();
                const uint32_t st = dst_lane + s_issue * STAGE_BYTES;
                unsigned char *stg = wbase + s_issue * STAGE_BYTES;
                if (cur.n < 0) {
                    if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(-1, 0, 0, 0);
                    done = true;
                } else {
                    const int nn = cur.n - half;
#pragma unroll
                    for (int i = 0; i < SR / 2; ++i) {
                        const unsigned j = (unsigned)__shfl_sync(FULL, cur.idx, 2 * i + half);
                        cp_async16(st + 2 * i * ROWB, src_lane + (size_t)j * 256, (2 * i < nn) ? 16 : 0);
                    }
                    if (lane < SR) reinterpret_cast<double *>(stg + W_OFF)[lane] = (lane < cur.n) ? (cur.w - p.mean_rating) * p.alpha : 0.0;
                    if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(cur.n, cur.item, cur.first, cur.last);
                }
                // the stage's "full" barrier completes when all 32 lanes' copies (and the plain stores above) have landed
                __threadfence_block();
                cp_async_mbar_arrive_noinc(wbase_s + s_issue * STAGE_BYTES + MBAR_OFF);
                s_issue = (s_issue + 1 == NS) ? 0 : s_issue + 1;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        return;
    }

    if (warp < NM) {
        // =============================================== MMA warp ===============================================
        const int gi = warp;
        unsigned char *wbase = rings + (size_t)gi * NS * STAGE_BYTES;
        int hint = (gi * NT) / NM;
        double c[10][2];
        double rrp[4];
#pragma unroll
        for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
        for (int a = 0; a < 4; ++a) rrp[a] = 0.0;
        const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);
        while (sflag[28 + gi] == 0) { }      // the producer has initialised the stage barriers
        __threadfence_block();
        unsigned phases = 0;
        int h = 0;
#pragma unroll 1
        for (;;) {
            mbar_wait(wbase_s + h * STAGE_BYTES + MBAR_OFF, (phases >> h) & 1u);
            phases ^= 1u << h;
            __syncwarp();
            unsigned char *stg = wbase + h * STAGE_BYTES;
            const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
            if (meta.x < 0) break;
            {
                const unsigned char *row = stg + t * ROWB + g * 8;
                const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
                if (meta.x == SR) {          // the common case, branch-free so the fragment loads run ahead of the DMMAs
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
            }
            __syncwarp();                     // every lane is done reading the stage
            if (lane == 0) mbar_arrive(bars + (gi * NS + h) * 8);
            h = (h + 1 == NS) ? 0 : h + 1;
            if (meta.w) {
                // ---- the item's Gram is complete: hand it to a tail warp
                int s = 0;
                if (lane == 0) {
                    s = hint;
                    for (;;) {
                        if (atomicCAS(const_cast<int *>(sstate) + s, 0, 1) == 0) break;
                        s = (s + 1 == NT) ? 0 : s + 1;
                        if (s == hint) __nanosleep(32);
                    }
                }
                s = __shfl_sync(FULL, s, 0);
                hint = (s + 1 == NT) ? 0 : s + 1;
                double *slot = reinterpret_cast<double *>(slots + (size_t)s * SLOT_BYTES) + lane;
#pragma unroll
                for (int b = 0; b < 10; ++b) {
                    slot[(2 * b) * 32] = c[b][0];
                    slot[(2 * b + 1) * 32] = c[b][1];
                    c[b][0] = 0.0; c[b][1] = 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) { slot[(20 + a) * 32] = rrp[a]; rrp[a] = 0.0; }
                if (lane == 0) *reinterpret_cast<int *>(slots + (size_t)s * SLOT_BYTES + SLOT_ROWS * 32 * 8) = meta.y;
                __syncwarp();
                if (lane == 0) { sstate[s] = 2; mbar_arrive(bars + 256 + s * 8); }
            }
        }
        if (lane == 0) {
            const int done = atomicAdd(const_cast<int *>(sstate) + NT, 1) + 1;
            if (done == NM)                   // the last MMA warp tells every tail warp to leave once its slot is empty
                for (int s = 0; s < NT; ++s) {
                    while (atomicCAS(const_cast<int *>(sstate) + s, 0, 3) != 0) __nanosleep(64);
                    mbar_arrive(bars + 256 + s * 8);
                }
        }
        return;
    }

    // ================================================= tail warp =================================================
    const int ti = warp - 2 * NM;
    const double *slot = reinterpret_cast<const double *>(slots + (size_t)ti * SLOT_BYTES) + lane;
    unsigned char *stg = scratch + (size_t)ti * TSCRATCH_BYTES;
    double *zy = reinterpret_cast<double *>(stg + V3_ZY_OFF), *zr = reinterpret_cast<double *>(stg + V3_ZR_OFF);
    double *wb = reinterpret_cast<double *>(stg + V3_B_OFF), *Lp = reinterpret_cast<double *>(stg);
#pragma unroll 1
    unsigned fphase = 0;
    for (;;) {
        mbar_wait(bars + 256 + ti * 8, fphase);
        fphase ^= 1u;
        const int st = sstate[ti];
        if (st == 3) break;
        double c[10][2];
        double rrp[4];
#pragma unroll
        for (int b = 0; b < 10; ++b) { c[b][0] = slot[(2 * b) * 32]; c[b][1] = slot[(2 * b + 1) * 32]; }
#pragma unroll
        for (int a = 0; a < 4; ++a) rrp[a] = slot[(20 + a) * 32];
        const int idx = *reinterpret_cast<const int *>(slots + (size_t)ti * SLOT_BYTES + SLOT_ROWS * 32 * 8);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sstate[ti] = 0; }   // the slot can take the next item while this one is solved
        if (DBG & 1) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 10; ++b) acc += c[b][0] + c[b][1];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc += rrp[a];
            p.items[(size_t)idx * 32 + lane] = acc;
            continue;
        }
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
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 1);
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) wb[8 * a + g] = srr0[8 * a + g] + rrp[a];
        }
        __syncwarp();
        const double z = (DBG & 8) ? 0.25 * lane : __dmul_rn(zy[lane], polar_mult(zr[lane]));
        double bb = wb[lane];
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    c[blk(I, J)][e] = fma(p.alpha, c[blk(I, J)][e], sLF[(8 * J + 2 * t + e) * LFS + 8 * I + g]);
        double myd = 1.0, myrinv = 1.0;
        bool ok = true;
        chol3_block_column<0>(c, myd, myrinv, lane, t);
        chol3_block_column<1>(c, myd, myrinv, lane, t);
        chol3_block_column<2>(c, myd, myrinv, lane, t);
        chol3_block_column<3>(c, myd, myrinv, lane, t);
        ok = __all_sync(FULL, myd > 0.0);
        const double myrs = rsqrt(myd);
        {
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * J + 2 * t + e;
                    const double rk = __shfl_sync(FULL, myrinv, k);
                    double *lq = Lp + (31 * k - ((k * (k - 1)) >> 1)) + g - k - 1;
#pragma unroll
                    for (int I = J; I < 4; ++I)
                        if (I > J || g > 2 * t + e) lq[8 * I] = c[blk(I, J)][e] * rk;
                }
        }
        __syncwarp();
        if (ok) {
            {
                const double *lf = Lp + lane - 1;
#pragma unroll
                for (int k = 0; k < 31; ++k) {
                    const double yk = __shfl_sync(FULL, bb, k);
                    if (lane > k) bb = fma(-lf[col_off1(k) - k], yk, bb);
                }
            }
            double yv = fma(bb, myrinv, myrs * z);
            {
                const double *lb = Lp + (31 * lane - ((lane * (lane - 1)) >> 1)) - lane - 1;
#pragma unroll
                for (int i = 31; i >= 1; --i) {
                    const double xi = __shfl_sync(FULL, yv, i);
                    if (lane < i) yv = fma(-lb[i], xi, yv);
                }
            }
            p.items[(size_t)idx * 32 + lane] = yv;
            for (int pr = 0; pr < p.npeers; ++pr) {
                double *dst = p.peers[pr];
                if (dst && dst != p.items) dst[(size_t)idx * 32 + lane] = yv;
            }
        } else if (lane == 0) {
            atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
        }
        __syncwarp();                         // the scratch is free again
    }
}


template <int NM, int NT, int NS, int DBG = 0>
cudaError_t launch_v7(bpmf_gpu_ctx *c, const StreamArgs &p, long long n)
{
    constexpr size_t smem = v7_smem_bytes<NM, NT, NS>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = items_stream32v7_kernel<NM, NT, NS, DBG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = c->sm_count;                              // persistent: one CTA per SM
    const long long need = (n + (long long)NM * CLAIM - 1) / ((long long)NM * CLAIM);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, (2 * NM + NT) * 32, smem, c->stream>>>(p);
    return cudaGetLastError();
}

// =====================================================================================================================
// Version 8: v3 with TWO items per warp in the tail. Under DMMA contention every dependent scalar fp64 instruction of
// a tail costs 15-56 cycles (bench_micro/fp64_latency.cu), and the tail is one dependent chain, so a warp that works on a
// single item leaves the pipe idle most of the time. Here a warp accumulates the Grams of two consecutive items (two
// accumulator sets, 80 registers) and then runs both tails interleaved instruction by instruction: two independent
// chains in flight per warp. 12 warps x 168 registers per SM; the tail scratch (packed Lu, with the normals / rhs vectors
// aliased on its first 768 bytes) is private to the warp, 2 x 3968 bytes, so the gather ring keeps running during tails.
// =====================================================================================================================
constexpr int V8_SCRATCH = LPACK1 * 8;              // per item; zy | zr | b alias the first 768 bytes (used before Lu exists)
template <int NS> constexpr __host__ __device__ int v8_warp_bytes() { return NS * STAGE_BYTES + 2 * V8_SCRATCH; }

template <int KB>
__device__ __forceinline__ void chol8_block_column(double (&c)[2][10][2], double (&myd)[2], double (&myrinv)[2], bool (&ok)[2], int lane, int t)
{
    constexpr int D = blk(KB, KB);
#pragma unroll 1
    for (int k2 = 0; k2 < 4; ++k2) {
        const int qsrc = (lane & ~3) | k2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = 2 * k2 + e;
            double p[2], bl0[2], bl1[2], a[2][4], rinv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                p[u] = __shfl_sync(FULL, c[u][D][e], 4 * k + k2);
                bl0[u] = __shfl_sync(FULL, c[u][D][e], 4 * (2 * t) + k2);
                bl1[u] = __shfl_sync(FULL, c[u][D][e], 4 * (2 * t + 1) + k2);
#pragma unroll
                for (int I = KB; I < 4; ++I) a[u][I] = __shfl_sync(FULL, c[u][blk(I, KB)][e], qsrc);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!(p[u] > 0.0)) ok[u] = false;
                rinv[u] = fast_rcp(p[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (lane == 8 * KB + k) { myd[u] = p[u]; myrinv[u] = rinv[u]; }
                bl0[u] = (2 * t > k) ? -(bl0[u] * rinv[u]) : 0.0;
                bl1[u] = (2 * t + 1 > k) ? -(bl1[u] * rinv[u]) : 0.0;
            }
#pragma unroll
            for (int I = KB; I < 4; ++I)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    c[u][blk(I, KB)][0] = fma(a[u][I], bl0[u], c[u][blk(I, KB)][0]);
                    c[u][blk(I, KB)][1] = fma(a[u][I], bl1[u], c[u][blk(I, KB)][1]);
                }
        }
    }
    if (KB < 3) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double fr[4][2], rv[2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) rv[kk] = __shfl_sync(FULL, myrinv[u], 8 * KB + 4 * kk + t);
#pragma unroll
            for (int I = KB + 1; I < 4; ++I)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int src = (lane & ~3) | (2 * kk + (t >> 1));
                    const double v0 = __shfl_sync(FULL, c[u][blk(I, KB)][0], src);
                    const double v1 = __shfl_sync(FULL, c[u][blk(I, KB)][1], src);
                    fr[I][kk] = (t & 1) ? v1 : v0;
                }
#pragma unroll
            for (int I = KB + 1; I < 4; ++I)
#pragma unroll
                for (int J = KB + 1; J <= I; ++J)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) dmma884(c[u][blk(I, J)][0], c[u][blk(I, J)][1], -fr[I][kk], fr[J][kk] * rv[kk]);
        }
    }
}

template <int NS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) items_stream32v8_kernel(StreamArgs p)
{
    constexpr int WARP_BYTES = v8_warp_bytes<NS>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    unsigned char *wbase = smem_raw + SHARED_BYTES + (size_t)warp * WARP_BYTES;
    const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);
    unsigned char *tscr = wbase + NS * STAGE_BYTES;                // 2 x V8_SCRATCH

    for (int e = tid; e < 1024; e += NW * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    // ---------------- fetch side: identical to v3
    int g_base = 0, g_n = 0, f_it = 0;
    int cpr = 0;
    int f_pos = 0, f_end = 0, f_start = 0, g_end = 0;
    const int32_t *g_idx = p.rowidx;
    const double *g_val = p.val;
    int32_t n_idx = 0;
    double n_w = 0.0;
    bool f_done = false;
    const unsigned char *src_lane = reinterpret_cast<const unsigned char *>(p.other) + (lane & 15) * 16;
    const uint32_t dst_lane = wbase_s + (lane >> 4) * ROWB + (lane & 15) * 16;
    const int half = lane >> 4;

    auto load_next = [&]() {
        const int q = f_pos + (lane & 15);
        n_idx = 0; n_w = 0.0;
        if (q < g_end) {
            n_idx = __ldg(g_idx + q);
            n_w = __ldg(g_val + q);
        }
    };
    auto claim = [&]() {
        int base = 0;
        if (lane == 0) base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.to) { f_done = true; return; }
        g_base = base;
        g_n = min(CLAIM, p.to - base);
        const int64_t c0 = __ldg(p.colptr + base);
        cpr = (lane <= g_n) ? (int)(__ldg(p.colptr + base + lane) - c0) : 0;
        g_idx = p.rowidx + c0;
        g_val = p.val + c0;
        f_it = 0;
        f_start = f_pos = 0;
        f_end = __shfl_sync(FULL, cpr, 1);
        g_end = __shfl_sync(FULL, cpr, g_n);
    };
    auto issue_stage = [&](int slot) {
        const uint32_t st = dst_lane + slot * STAGE_BYTES;
        unsigned char *stg = wbase + slot * STAGE_BYTES;
        if (f_done) {
            if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(-1, 0, 0, 0);
            cp_async_commit();
            return;
        }
        const int n = min(SR, f_end - f_pos);
        const int nn = n - half;
#pragma unroll
        for (int i = 0; i < SR / 2; ++i) {
            const unsigned j = (unsigned)__shfl_sync(FULL, n_idx, 2 * i + half);
            cp_async16(st + 2 * i * ROWB, src_lane + (size_t)j * 256, (2 * i < nn) ? 16 : 0);
        }
        if (lane < SR) reinterpret_cast<double *>(stg + W_OFF)[lane] = (lane < n) ? (n_w - p.mean_rating) * p.alpha : 0.0;
        const int last = (f_pos + n == f_end);
        if (lane == 0) *reinterpret_cast<int4 *>(stg + META_OFF) = make_int4(n, g_base + f_it, f_pos == f_start, last);
        cp_async_commit();
        f_pos += n;
        if (last) {
            ++f_it;
            if (f_it >= g_n) claim();
            else { f_start = f_end; f_end = __shfl_sync(FULL, cpr, f_it + 1); }
        }
        if (!f_done) load_next();
    };

    claim();
    if (!f_done) load_next();
#pragma unroll 1
    for (int s = 0; s < NS; ++s) issue_stage(s);

    int h = 0;
    // Gram of the next item of the stream into (cc, rr); returns its index, or -1 at the end of the stream
    auto run_gram = [&](double (&cc)[10][2], double (&rr)[4]) -> int {
#pragma unroll
        for (int b = 0; b < 10; ++b) { cc[b][0] = 0.0; cc[b][1] = 0.0; }
#pragma unroll
        for (int a = 0; a < 4; ++a) rr[a] = 0.0;
#pragma unroll 1
        for (;;) {
            cp_async_wait<NS - 1>();
            __syncwarp();
            unsigned char *stg = wbase + h * STAGE_BYTES;
            const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
            if (meta.x < 0) return -1;
            const unsigned char *row = stg + t * ROWB + g * 8;
            const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
            if (meta.x > 0) gram_group(cc, rr, row, wq);
            if (meta.x > 4) gram_group(cc, rr, row + 4 * ROWB, wq + 4);
            if (meta.x > 8) gram_group(cc, rr, row + 8 * ROWB, wq + 8);
            if (meta.x > 12) gram_group(cc, rr, row + 12 * ROWB, wq + 12);
            __syncwarp();                 // every lane is done reading slot h
            issue_stage(h);
            h = (h + 1 == NS) ? 0 : h + 1;
            if (meta.w) return meta.y;
        }
    };

    double c[2][10][2];
    double rrp[2][4];
#pragma unroll 1
    for (;;) {
        int idx[2];
        idx[0] = run_gram(c[0], rrp[0]);
        if (idx[0] < 0) break;
        idx[1] = run_gram(c[1], rrp[1]);
        const bool two = idx[1] >= 0;     // odd item at the end of the stream: the second tail runs on MM = LambdaF and is dropped
        // ---------------- both tails, interleaved ----------------
        double z[2], bb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double *zy = reinterpret_cast<double *>(tscr + u * V8_SCRATCH), *zr = zy + 32, *wb = zy + 64;
            const uint32_t seed = (uint32_t)(((long long)(two || u == 0 ? idx[u] : 0) + 1) * 32ll * ((long long)p.iter + 1));
            int have = 0;
            for (uint32_t base = 0; have < 32; base += 32) {
                const U4 bk = stream_block(seed, base + lane);
                const Polar pa = polar_attempt(bk.v[3], bk.v[2], bk.v[1], bk.v[0]);
                const unsigned m = __ballot_sync(FULL, pa.ok);
                const int n = have + __popc(m & ((1u << lane) - 1u));
                if (pa.ok && n < 32) { zy[n] = pa.y; zr[n] = pa.r2; }
                have += __popc(m);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                rrp[u][a] += __shfl_xor_sync(FULL, rrp[u][a], 1);
                rrp[u][a] += __shfl_xor_sync(FULL, rrp[u][a], 2);
            }
            if (t == 0) {
#pragma unroll
                for (int a = 0; a < 4; ++a) wb[8 * a + g] = srr0[8 * a + g] + rrp[u][a];
            }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double *zy = reinterpret_cast<const double *>(tscr + u * V8_SCRATCH), *zr = zy + 32, *wb = zy + 64;
            z[u] = __dmul_rn(zy[lane], polar_mult(zr[lane]));
            bb[u] = wb[lane];
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        c[u][blk(I, J)][e] = fma(p.alpha, c[u][blk(I, J)][e], sLF[(8 * J + 2 * t + e) * LFS + 8 * I + g]);
        }
        double myd[2] = {1.0, 1.0}, myrinv[2] = {1.0, 1.0};
        bool ok[2] = {true, true};
        chol8_block_column<0>(c, myd, myrinv, ok, lane, t);
        chol8_block_column<1>(c, myd, myrinv, ok, lane, t);
        chol8_block_column<2>(c, myd, myrinv, ok, lane, t);
        chol8_block_column<3>(c, myd, myrinv, ok, lane, t);
        double myrs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) myrs[u] = rsqrt(myd[u]);
        __syncwarp();                     // zy / zr / b have been read by every lane: Lu may overwrite them
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double *Lp = reinterpret_cast<double *>(tscr + u * V8_SCRATCH);
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * J + 2 * t + e;
                    const double rk = __shfl_sync(FULL, myrinv[u], k);
                    double *lq = Lp + (31 * k - ((k * (k - 1)) >> 1)) + g - k - 1;
#pragma unroll
                    for (int I = J; I < 4; ++I)
                        if (I > J || g > 2 * t + e) lq[8 * I] = c[u][blk(I, J)][e] * rk;
                }
        }
        __syncwarp();
        {
            const double *lf0 = reinterpret_cast<const double *>(tscr) + lane - 1;
            const double *lf1 = reinterpret_cast<const double *>(tscr + V8_SCRATCH) + lane - 1;
#pragma unroll
            for (int k = 0; k < 31; ++k) {
                const double y0 = __shfl_sync(FULL, bb[0], k);
                const double y1 = __shfl_sync(FULL, bb[1], k);
                if (lane > k) {
                    bb[0] = fma(-lf0[col_off1(k) - k], y0, bb[0]);
                    bb[1] = fma(-lf1[col_off1(k) - k], y1, bb[1]);
                }
            }
        }
        double yv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) yv[u] = fma(bb[u], myrinv[u], myrs[u] * z[u]);
        {
            const int off = (31 * lane - ((lane * (lane - 1)) >> 1)) - lane - 1;
            const double *lb0 = reinterpret_cast<const double *>(tscr) + off;
            const double *lb1 = reinterpret_cast<const double *>(tscr + V8_SCRATCH) + off;
#pragma unroll
            for (int i = 31; i >= 1; --i) {
                const double x0 = __shfl_sync(FULL, yv[0], i);
                const double x1 = __shfl_sync(FULL, yv[1], i);
                if (lane < i) {
                    yv[0] = fma(-lb0[i], x0, yv[0]);
                    yv[1] = fma(-lb1[i], x1, yv[1]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            if (ok[u]) {
                p.items[(size_t)idx[u] * 32 + lane] = yv[u];
                for (int pr = 0; pr < p.npeers; ++pr) {
                    double *dst = p.peers[pr];
                    if (dst && dst != p.items) dst[(size_t)idx[u] * 32 + lane] = yv[u];
                }
            } else if (lane == 0) {
                atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx[u]);
            }
        }
        __syncwarp();                     // the scratch is free again
        if (!two) break;
    }
    cp_async_wait<0>();
}

template <int NS, int NW>
cudaError_t launch_v8(bpmf_gpu_ctx *c, const StreamArgs &p, long long n)
{
    constexpr size_t smem = (size_t)NW * v8_warp_bytes<NS>() + SHARED_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = items_stream32v8_kernel<NS, NW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = c->sm_count;                              // persistent: one CTA per SM
    const long long need = (n + (long long)NW * CLAIM - 1) / ((long long)NW * CLAIM);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, NW * 32, smem, c->stream>>>(p);
    return cudaGetLastError();
}

