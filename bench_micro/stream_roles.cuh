// stream_roles.cuh — EXPERIMENT (measured slower, profiles/r02_tune_warp_roles.log): the K = 32 stream kernel with warp ROLES.
// Included by bpmf_b200/csrc/stream_kernel.cu only when the library is built with BPMF_STREAM_PROBES=1; selected with
// bpmf_gpu_debug_set_roles (bench_micro/tune_roles.py). Not part of the product build.
#pragma once

// =====================================================================================================================
// Version 5: warp ROLES. DMMA and scalar fp64 share one pipe per scheduler and the arbiter serves warps by instruction:
// a dependent DFMA of one warp waits ~32 cycles per warp that streams DMMAs on the same scheduler
// (profiles/r02_fp64_mix_microbench*.txt: 9 cycles alone, 42 / 74 / 106 with one / two / three streaming warps), while
// the pipe TIME of the two kinds simply adds. In v3 every warp alternates between its Gram and its tail, so a tail (a
// chain of ~260 dependent fp64 instructions) usually runs against one to three Gram streams. Here the roles are split:
//   Gram warps  (bit w of cfg.gram_mask) run the fetch ring + the DMMA Gram of v3 and, when an item is complete, hand its
//               accumulators (24 doubles per lane, the DMMA layout as it is) to a transit slot in shared memory;
//   tail warps  take a full slot into registers, free it, and run tail32_warp (normals, LDL^T, solves, store).
// Which warps are which decides what a tail competes with: warp w issues on scheduler w % 4, so e.g. mask 0x0000f is one
// Gram stream per scheduler with four tail warps each, 0x77777 keeps scheduler 3 free of DMMAs altogether.
// Slots: state 0 free -> 2 being written -> 1 full -> 3 being read -> 0. Gram warps that run out of items count themselves
// in `done`; a tail warp leaves when it saw done == number of Gram warps BEFORE a scan that found nothing.
// =====================================================================================================================
struct V5Cfg {
    unsigned gram_mask;   // bit w: warp w is a Gram warp
    int ns;               // ring stages per Gram warp
    int nslot;            // transit slots
};
constexpr int V5_SLOT_DOUBLES = 24 * 32;
constexpr int V5_SLOT_BYTES = V5_SLOT_DOUBLES * 8 + 32;      // + item index
constexpr int V5_CTRL_BYTES = 256;                           // slot states (<= 48) + done counter
constexpr int V5_TSCRATCH = ((V3_B_OFF + 256 + 15) / 16) * 16;
inline size_t v5_smem_bytes(const V5Cfg &cfg, int nw)
{
    const int ng = __builtin_popcount(cfg.gram_mask & ((nw >= 32) ? 0xffffffffu : ((1u << nw) - 1u)));
    return (size_t)SHARED_BYTES + (size_t)cfg.nslot * V5_SLOT_BYTES + V5_CTRL_BYTES + (size_t)ng * cfg.ns * STAGE_BYTES + (size_t)(nw - ng) * V5_TSCRATCH;
}

__device__ __forceinline__ void cp_async_wait_dyn(int pending)
{
    switch (pending) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
    }
}

template <int NW, int DBG>
__global__ void __launch_bounds__(NW * 32, 1) items_stream32v5_kernel(StreamArgs p, V5Cfg cfg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *sLF = reinterpret_cast<double *>(smem_raw);            // LambdaF(i,k) at sLF[k * LFS + i]
    double *srr0 = sLF + 32 * LFS;                                 // LambdaF * mu
    unsigned char *slots = smem_raw + SHARED_BYTES;
    volatile int *state = reinterpret_cast<volatile int *>(slots + (size_t)cfg.nslot * V5_SLOT_BYTES);
    int *state_nv = const_cast<int *>(state);
    volatile int *done = state + 48;
    unsigned char *regions = slots + (size_t)cfg.nslot * V5_SLOT_BYTES + V5_CTRL_BYTES;
    const unsigned wmask = (NW >= 32) ? 0xffffffffu : ((1u << NW) - 1u);
    const unsigned gmask = cfg.gram_mask & wmask;
    const int n_gram = __popc(gmask);
    const bool is_gram = (gmask >> warp) & 1u;
    const int my_rank = __popc(gmask & ((1u << warp) - 1u));       // rank among the Gram warps (for a Gram warp)
    const int NS = cfg.ns;

    for (int e = tid; e < 1024; e += NW * 32) sLF[(e >> 5) * LFS + (e & 31)] = p.LambdaF[e];
    if (tid < 64) state_nv[tid] = 0;
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[j * LFS + tid] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    double c[10][2];
    double rrp[4];

    if (!is_gram) {
        // ------------------------------------------------------------------------------------------------ tail warp
        unsigned char *scratch = regions + (size_t)n_gram * NS * STAGE_BYTES + (size_t)(warp - my_rank) * V5_TSCRATCH;
        int start = warp % cfg.nslot;
        for (;;) {
            int s = -1;
            if (lane == 0) {
                for (;;) {
                    const int d = *done;
                    for (int k = 0; k < cfg.nslot; ++k) {
                        int q = start + k; if (q >= cfg.nslot) q -= cfg.nslot;
                        if (state[q] == 1 && atomicCAS(state_nv + q, 1, 3) == 1) { s = q; break; }
                    }
                    if (s >= 0) break;
                    if (d == n_gram) { s = -2; break; }
                    __nanosleep(100);
                }
            }
            s = __shfl_sync(FULL, s, 0);
            if (s < 0) break;
            __threadfence_block();
            const double *sl = reinterpret_cast<const double *>(slots + (size_t)s * V5_SLOT_BYTES) + lane;
#pragma unroll
            for (int b = 0; b < 10; ++b) { c[b][0] = sl[(2 * b) * 32]; c[b][1] = sl[(2 * b + 1) * 32]; }
#pragma unroll
            for (int a = 0; a < 4; ++a) rrp[a] = sl[(20 + a) * 32];
            const int idx = *reinterpret_cast<const int *>(slots + (size_t)s * V5_SLOT_BYTES + V5_SLOT_DOUBLES * 8);
            __syncwarp();
            if (lane == 0) { __threadfence_block(); state[s] = 0; }
            start = s + 1; if (start >= cfg.nslot) start = 0;
            tail32_warp<DBG, false>(c, rrp, idx, scratch, sLF, srr0, p, lane);
            __syncwarp();
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------------- Gram warp
    unsigned char *wbase = regions + (size_t)my_rank * NS * STAGE_BYTES;
    const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(wbase);
    int g_base = 0, g_n = 0, f_it = 0;
    int cpr = 0;                          // per lane: colptr[g_base + lane] - colptr[g_base]
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
        int base = 0, lim = p.bulk_end;
        if (lane == 0) {
            base = p.from + (int)atomicAdd(p.work_counter, (unsigned)CLAIM);
            if (base >= p.bulk_end) {
                base = p.bulk_end + (int)atomicAdd(p.work_counter + 1, (unsigned)CLAIM_TAIL);
                lim = min(p.to, base + CLAIM_TAIL);
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
        f_pos = f_start;
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
#pragma unroll
    for (int b = 0; b < 10; ++b) { c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; ++a) rrp[a] = 0.0;

    int h = 0, sstart = my_rank % cfg.nslot;
#pragma unroll 1
    for (;;) {
        cp_async_wait_dyn(NS - 1);
        __syncwarp();
        unsigned char *stg = wbase + h * STAGE_BYTES;
        const int4 meta = *reinterpret_cast<const int4 *>(stg + META_OFF);
        if (meta.x < 0) break;
        {
            const unsigned char *row = stg + t * ROWB + g * 8;
            const double *wq = reinterpret_cast<const double *>(stg + W_OFF) + t;
            if (!(DBG & 2)) {
                if (meta.x == SR) {
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
        }
        __syncwarp();                     // every lane is done reading slot h
        issue_stage(h);                   // refill first: the hand-off below may have to wait for a free slot
        h = (h + 1 == NS) ? 0 : h + 1;
        if (!meta.w) continue;
        // ---------------- the item's Gram is complete: hand the accumulators to a tail warp
        int s = 0;
        if (lane == 0) {
            for (;;) {
                s = -1;
                for (int k = 0; k < cfg.nslot; ++k) {
                    int q = sstart + k; if (q >= cfg.nslot) q -= cfg.nslot;
                    if (state[q] == 0 && atomicCAS(state_nv + q, 0, 2) == 0) { s = q; break; }
                }
                if (s >= 0) break;
                __nanosleep(64);
            }
        }
        s = __shfl_sync(FULL, s, 0);
        sstart = s + 1; if (sstart >= cfg.nslot) sstart = 0;
        double *sl = reinterpret_cast<double *>(slots + (size_t)s * V5_SLOT_BYTES) + lane;
#pragma unroll
        for (int b = 0; b < 10; ++b) { sl[(2 * b) * 32] = c[b][0]; sl[(2 * b + 1) * 32] = c[b][1]; c[b][0] = 0.0; c[b][1] = 0.0; }
#pragma unroll
        for (int a = 0; a < 4; ++a) { sl[(20 + a) * 32] = rrp[a]; rrp[a] = 0.0; }
        if (lane == 0) *reinterpret_cast<int *>(slots + (size_t)s * V5_SLOT_BYTES + V5_SLOT_DOUBLES * 8) = meta.y;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); state[s] = 1; }
    }
    cp_async_wait<0>();
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicAdd(const_cast<int *>(done), 1); }
}

template <int NW>
static cudaError_t launch_v5(bpmf_gpu_ctx *c, const StreamArgs &p, long long n, const V5Cfg &cfg)
{
    const size_t smem = v5_smem_bytes(cfg, NW);
    if (smem > 227 * 1024 || cfg.nslot < 1 || cfg.nslot > 48 || cfg.ns < 1 || cfg.ns > 8) return cudaErrorInvalidConfiguration;
    const unsigned gm = cfg.gram_mask & ((1u << NW) - 1u);
    if (gm == 0 || gm == ((1u << NW) - 1u)) return cudaErrorInvalidConfiguration;   // both roles must be present
    auto kern = items_stream32v5_kernel<NW, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = c->sm_count;
    const long long ng = __builtin_popcount(gm);
    const long long need = (n + ng * CLAIM_TAIL - 1) / (ng * CLAIM_TAIL);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, NW * 32, smem, c->stream>>>(p, cfg);
    return cudaGetLastError();
}

