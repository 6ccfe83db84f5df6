// capi.cu — the extern "C" surface declared in include/bpmf_gpu.h. Host-side only: owns device memory,
// orders the kernels of one Sys::sample (c++/sample.cpp:341-385) on the context's stream and turns
// CUDA / kernel errors into return codes. There is deliberately no CPU fallback anywhere in this file.
#include "../../include/bpmf_gpu.h"
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

using namespace bpmf;

namespace {

std::string g_create_err;

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            char b__[512];                                                                             \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            ctx->err = b__;                                                                            \
            return BPMF_GPU_ECUDA;                                                                     \
        }                                                                                              \
    } while (0)

int fail(bpmf_gpu_ctx *ctx, int code, const std::string &msg)
{
    ctx->err = msg;
    return code;
}

template <typename T>
void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

void free_side(SideDev &s)
{
    dfree(s.colptr); dfree(s.rowidx); dfree(s.val); dfree(s.wval);
    dfree(s.t_colptr); dfree(s.t_rowidx); dfree(s.t_col); dfree(s.t_val); dfree(s.pavg); dfree(s.pm2);
    dfree(s.items_own); s.items = nullptr;
    dfree(s.peers_dev); dfree(s.stat_peers_dev);
    dfree(s.hp.mu); dfree(s.hp.LambdaU); dfree(s.hp.LambdaF);
    dfree(s.hp_next.mu); dfree(s.hp_next.LambdaU); dfree(s.hp_next.LambdaF);
    dfree(s.sum); dfree(s.prod); dfree(s.cov); dfree(s.norm); dfree(s.partials); dfree(s.pred_partials);
    dfree(s.work_counter); dfree(s.aggrMu); dfree(s.aggrLambda); dfree(s.propLambda);
    dfree(s.hv_item); dfree(s.hv_first); dfree(s.hv_p0); dfree(s.hv_p1); dfree(s.hv_partials);
    s = SideDev();
}

bool side_ok(int side) { return side == 0 || side == 1; }

// deferred kernel-side errors (the word kernels atomicMax into). It is read AND reset by one atomic exchange on the device
// (launch_fetch_error): a hyper draw running ahead on the auxiliary stream may report into the word at any time, and a
// read followed by a separate clear could wipe what it wrote in between.
int check_device_error(bpmf_gpu_ctx *ctx)
{
    CU(launch_fetch_error(ctx));
    CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const unsigned long long w = *ctx->h_err;
    if (!w) return BPMF_GPU_OK;
    const unsigned code = (unsigned)(w >> 32), detail = (unsigned)(w & 0xffffffffu);
    char b[256];
    if (code == 3) { snprintf(b, sizeof b, "Cholesky failed (item %u)", detail); return fail(ctx, BPMF_GPU_ECHOLESKY, b); }
    if (code == 4) return fail(ctx, BPMF_GPU_ERNG, "hyper-parameter draw ran out of pre-generated Philox blocks");
    if (code == 5) { snprintf(b, sizeof b, "hyper-parameter draw: T_c not positive definite (pivot %u)", detail); return fail(ctx, BPMF_GPU_ECHOLESKY, b); }
    if (code == 6) { snprintf(b, sizeof b, "cross-GPU barrier: rank %u did not arrive", detail); return fail(ctx, BPMF_GPU_ECUDA, b); }
    snprintf(b, sizeof b, "device error word %llx", w);
    return fail(ctx, BPMF_GPU_ECUDA, b);
}

// per-item prior precisions (propagated posterior) are read by the PROP instantiations of the K = 32 stream kernel and of
// its heavy-item tail, by the CTA-per-item kernel and by the any-K kernel
static bool prop_on_stream(const bpmf_gpu_ctx *ctx, int side) { (void)side; return ctx->K == 32; }

int pick_variant(const bpmf_gpu_ctx *ctx, int side, int v)
{
    // per-item prior precisions (propagated posterior) are read by the any-K kernel and, for K = 32 on a side without
    // heavy items, by the PROP instantiation of the stream kernel
    if (v == BPMF_GPU_KERNEL_AUTO) {
        if (ctx->side[side].propLambda)
            return prop_on_stream(ctx, side) ? BPMF_GPU_KERNEL_STREAM : block_kernel_supports(ctx->K) ? BPMF_GPU_KERNEL_BLOCK : BPMF_GPU_KERNEL_EXACT;
        if (ctx->K == 32) return BPMF_GPU_KERNEL_STREAM;
        return block_kernel_supports(ctx->K) ? BPMF_GPU_KERNEL_BLOCK : BPMF_GPU_KERNEL_EXACT;
    }
    return v;
}

}  // namespace

extern "C" {

const char *bpmf_gpu_last_error(const bpmf_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bpmf_gpu_create(bpmf_gpu_ctx **out, int device, int num_latent)
{
    if (!out) return BPMF_GPU_EINVAL;
    *out = nullptr;
    if (num_latent < 1 || num_latent > 128) { g_create_err = "num_latent must be in [1,128]"; return BPMF_GPU_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libbpmf_b200 has no CPU fallback)";
        return BPMF_GPU_ENODEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return BPMF_GPU_EINVAL; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return BPMF_GPU_ECUDA; }
    if (prop.major != 10) {
        g_create_err = std::string("device ") + prop.name + " is not sm_100 (this library is built for sm_100a only)";
        return BPMF_GPU_ENODEVICE;
    }
    bpmf_gpu_ctx *ctx = new (std::nothrow) bpmf_gpu_ctx();
    if (!ctx) return BPMF_GPU_EINVAL;
    ctx->device = device; ctx->K = num_latent; ctx->sm_count = prop.multiProcessorCount;
    const int K = num_latent, KK = K * K;
    auto bail = [&](cudaError_t ee) { g_create_err = cudaGetErrorString(ee); bpmf_gpu_destroy(ctx); return BPMF_GPU_ECUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_err, 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e);   // [0] the word, [1] what was fetched
    if ((e = cudaMemset(ctx->d_err, 0, 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e);
    if ((e = cudaMallocHost(&ctx->h_err, sizeof(unsigned long long))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_zero_row, sizeof(double) * 128)) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->d_zero_row, 0, sizeof(double) * 128)) != cudaSuccess) return bail(e);
    if ((e = cudaMallocHost(&ctx->h_pinned, sizeof(double) * (KK + K + 8))) != cudaSuccess) return bail(e);
    for (int i = 0; i < bpmf_gpu_ctx::EV_RING; ++i) {
        if ((e = cudaEventCreate(&ctx->ev0[i])) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreate(&ctx->ev1[i])) != cudaSuccess) return bail(e);
    }
    for (int sd = 0; sd < 2; ++sd) {
        HyperScratch &h = ctx->hs[sd];
        // expected consumption is 1.27 (K^2+K) + K/2 blocks; 2 (K^2+4K) + 64 is > 40 standard deviations above it
        h.nblk = 2 * (KK + 4 * K) + 64;
        if ((e = cudaMalloc(&h.words, sizeof(uint32_t) * 4 * h.nblk)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.acc, 2 * h.nblk)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.rank, sizeof(int) * 2 * h.nblk)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.pos_of_rank, sizeof(int) * 2 * h.nblk)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.row_start, sizeof(int) * (K + 1))) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.row_cls, sizeof(int) * (K + 1))) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.mats, sizeof(double) * 6 * KK)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.vecs, sizeof(double) * 4 * K)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.piv, sizeof(int) * K)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc(&h.host_in, sizeof(double) * (KK + K))) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_stats[sd], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_hyper[sd], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_sdone[sd], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_items[sd], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    for (int i = 0; i < bpmf_gpu_ctx::HOST_PARTS; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_part[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming)) != cudaSuccess) return bail(e);
    ctx->overlap_hyper = getenv("BPMF_NO_HYPER_OVERLAP") == nullptr;
    // BPMF_STATS_MAIN: the reductions of a sweep on the main stream behind its item kernel (rounds 1 - 2a), all SMs to the item kernels
    // (K == 32 only: the any-K partial kernel is a wide grid that would take the SMs of the next item kernel in the gap between two)
    ctx->stats_aux = ctx->overlap_hyper && num_latent == 32 && getenv("BPMF_STATS_MAIN") == nullptr;
    if (const char *r = getenv("BPMF_RESERVE_SMS")) ctx->reserve_sms = atoi(r) < 0 ? 0 : atoi(r);
    *out = ctx;
    return BPMF_GPU_OK;
}

int bpmf_gpu_destroy(bpmf_gpu_ctx *ctx)
{
    if (!ctx) return BPMF_GPU_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_side(ctx->side[0]); free_side(ctx->side[1]);
    for (void *m : ctx->ipc_mapped) cudaIpcCloseMemHandle(m);
    ctx->ipc_mapped.clear();
    for (int sd = 0; sd < 2; ++sd) {
        HyperScratch &h = ctx->hs[sd];
        dfree(h.words); dfree(h.acc); dfree(h.rank); dfree(h.pos_of_rank); dfree(h.row_start); dfree(h.row_cls);
        dfree(h.mats); dfree(h.vecs); dfree(h.piv); dfree(h.host_in);
        if (ctx->ev_stats[sd]) cudaEventDestroy(ctx->ev_stats[sd]);
        if (ctx->ev_hyper[sd]) cudaEventDestroy(ctx->ev_hyper[sd]);
        if (ctx->ev_sdone[sd]) cudaEventDestroy(ctx->ev_sdone[sd]);
        if (ctx->ev_items[sd]) cudaEventDestroy(ctx->ev_items[sd]);
    }
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < bpmf_gpu_ctx::HOST_PARTS; ++i)
        if (ctx->ev_part[i]) cudaEventDestroy(ctx->ev_part[i]);
    if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
    dfree(ctx->d_err); dfree(ctx->d_zero_row);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int i = 0; i < bpmf_gpu_ctx::EV_RING; ++i) {
        if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
    }
    delete ctx;
    return BPMF_GPU_OK;
}

int bpmf_gpu_num_latent(const bpmf_gpu_ctx *ctx) { return ctx ? ctx->K : -1; }

int bpmf_gpu_set_stream(bpmf_gpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return BPMF_GPU_EINVAL;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return BPMF_GPU_OK;
}

int bpmf_gpu_sync(bpmf_gpu_ctx *ctx)
{
    if (!ctx) return BPMF_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->aux_stream));   // the reductions / next hyper draw of the last sweeps
    return check_device_error(ctx);
}

// Host-driven changes of a context's device state (latent matrices, ranges, peers, ...) wait for the reductions that may
// still be reading it on the auxiliary stream.
static int quiesce_aux(bpmf_gpu_ctx *ctx)
{
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->aux_stream));
    return BPMF_GPU_OK;
}

static int setup_side(bpmf_gpu_ctx *ctx, SideDev &s, const int64_t *colptr);
static int finish_test(bpmf_gpu_ctx *ctx, SideDev &s);

int bpmf_gpu_load_side(bpmf_gpu_ctx *ctx, int side, int num_items, int num_other, const int64_t *colptr, const int32_t *rowidx,
                       const double *val, double mean_rating)
{
    if (!ctx || !side_ok(side) || num_items < 0 || num_other < 0 || !colptr) return BPMF_GPU_EINVAL;
    const int64_t nnz = colptr[num_items];
    if (colptr[0] != 0 || nnz < 0 || (nnz > 0 && (!rowidx || !val))) return fail(ctx, BPMF_GPU_EINVAL, "bad CSC arrays");
    for (int i = 0; i < num_items; ++i)
        if (colptr[i + 1] < colptr[i]) return fail(ctx, BPMF_GPU_EINVAL, "colptr not monotone");
    for (int64_t p = 0; p < nnz; ++p)
        if (rowidx[p] < 0 || rowidx[p] >= num_other) return fail(ctx, BPMF_GPU_EINVAL, "row index out of range");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->aux_stream));
    SideDev &s = ctx->side[side];
    free_side(s);
    s.num = num_items; s.num_other = num_other; s.nnz = nnz;
    s.mean_rating = mean_rating;
    CU(cudaMalloc(&s.colptr, sizeof(int64_t) * ((size_t)num_items + 1)));
    CU(cudaMalloc(&s.rowidx, sizeof(int32_t) * (size_t)(nnz + 32)));  // +32: the DMMA kernel reads index chunks of 32
    CU(cudaMalloc(&s.val, sizeof(double) * (size_t)(nnz + 32)));
    CU(cudaMemset(s.rowidx, 0, sizeof(int32_t) * (size_t)(nnz + 32)));
    CU(cudaMemset(s.val, 0, sizeof(double) * (size_t)(nnz + 32)));
    CU(cudaMemcpy(s.colptr, colptr, sizeof(int64_t) * ((size_t)num_items + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        CU(cudaMemcpy(s.rowidx, rowidx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s.val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    return setup_side(ctx, s, colptr);
}

int bpmf_gpu_load_side_slice(bpmf_gpu_ctx *ctx, int side, int num_items, int num_other, int from, int to, const int64_t *colptr_slice,
                             const int32_t *rowidx_slice, const double *val_slice, double mean_rating)
{
    if (!ctx || !side_ok(side) || num_items < 0 || num_other < 0 || from < 0 || to < from || to > num_items || !colptr_slice) return BPMF_GPU_EINVAL;
    const int n = to - from;
    const int64_t nnz = colptr_slice[n];
    if (colptr_slice[0] != 0 || nnz < 0 || (nnz > 0 && (!rowidx_slice || !val_slice))) return fail(ctx, BPMF_GPU_EINVAL, "bad CSC arrays");
    // the kernels index items and ratings globally: a full-length column pointer array in which every item outside the
    // slice is empty costs 8 bytes per item and keeps them unchanged
    std::vector<int64_t> colptr((size_t)num_items + 1);
    for (int i = 0; i <= num_items; ++i) colptr[(size_t)i] = i < from ? 0 : (i > to ? nnz : colptr_slice[i - from]);
    for (int i = 0; i < n; ++i)
        if (colptr_slice[i + 1] < colptr_slice[i]) return fail(ctx, BPMF_GPU_EINVAL, "colptr not monotone");
    const int rc = bpmf_gpu_load_side(ctx, side, num_items, num_other, colptr.data(), rowidx_slice, val_slice, mean_rating);
    if (rc) return rc;
    SideDev &s = ctx->side[side];
    s.slice_from = from; s.slice_to = to;
    s.from = from; s.to = to;
    return BPMF_GPU_OK;
}

// Everything of a side but the compressed matrix itself: s.colptr / rowidx / val are on the device, s.num, num_other, nnz
// and mean_rating are set; colptr is the host copy of the column pointers.
static int setup_side(bpmf_gpu_ctx *ctx, SideDev &s, const int64_t *colptr)
{
    const int K = ctx->K, KK = K * K;
    const int num_items = s.num;
    const int64_t nnz = s.nnz;
    s.from = 0; s.to = num_items; s.iter = -1;
    s.slice_from = 0; s.slice_to = num_items;
    const size_t nitems = (size_t)K * (size_t)(num_items > 0 ? num_items : 1);
    CU(cudaMalloc(&s.items_own, sizeof(double) * nitems));
    CU(cudaMemset(s.items_own, 0, sizeof(double) * nitems));  // items().setZero() (sample.cpp:185)
    s.items = s.items_own;
    CU(cudaMalloc(&s.peers_dev, sizeof(double *) * MAX_PEERS));
    CU(cudaMemset(s.peers_dev, 0, sizeof(double *) * MAX_PEERS));
    CU(cudaMalloc(&s.hp.mu, sizeof(double) * K));
    CU(cudaMalloc(&s.hp.LambdaU, sizeof(double) * KK));
    CU(cudaMalloc(&s.hp.LambdaF, sizeof(double) * KK));
    CU(cudaMemset(s.hp.mu, 0, sizeof(double) * K));
    CU(cudaMemset(s.hp.LambdaU, 0, sizeof(double) * KK));
    CU(cudaMemset(s.hp.LambdaF, 0, sizeof(double) * KK));
    CU(cudaMalloc(&s.hp_next.mu, sizeof(double) * K));
    CU(cudaMalloc(&s.hp_next.LambdaU, sizeof(double) * KK));
    CU(cudaMalloc(&s.hp_next.LambdaF, sizeof(double) * KK));
    CU(cudaMalloc(&s.sum, sizeof(double) * K));
    CU(cudaMalloc(&s.prod, sizeof(double) * KK));
    CU(cudaMalloc(&s.cov, sizeof(double) * KK));
    CU(cudaMalloc(&s.norm, sizeof(double)));
    CU(cudaMemset(s.sum, 0, sizeof(double) * K));       // sum.setZero(); cov.setZero(); norm = 0 (sample.cpp:187-189)
    CU(cudaMemset(s.prod, 0, sizeof(double) * KK));
    CU(cudaMemset(s.cov, 0, sizeof(double) * KK));
    CU(cudaMemset(s.norm, 0, sizeof(double)));
    // + 2 x MAX_PEERS arrival words of the cross-GPU barriers (peer_barrier_kernel, one set per kind) behind the partials
    CU(cudaMalloc(&s.partials, sizeof(double) * ((size_t)STATS_BLOCKS * (KK + K + 1) + 2 * MAX_PEERS)));
    CU(cudaMemset(s.partials, 0, sizeof(double) * ((size_t)STATS_BLOCKS * (KK + K + 1) + 2 * MAX_PEERS)));
    CU(cudaMalloc(&s.stat_peers_dev, sizeof(double *) * MAX_PEERS));
    CU(cudaMemset(s.stat_peers_dev, 0, sizeof(double *) * MAX_PEERS));
    CU(cudaMalloc(&s.work_counter, 2 * sizeof(unsigned int)));
    CU(cudaMemset(s.work_counter, 0, 2 * sizeof(unsigned int)));
    // skew handling for the K == 32 stream kernel: items far heavier than the rest are sampled by the chunked path
    if ((K == 32 || block_kernel_supports(K)) && num_items > 0) {
        // heavy = more ratings than max(threshold, 16 x the mean); the bar is doubled until at most MAX_HEAVY items and
        // MAX_CHUNKS chunks (6 KB of partial Gram each at K = 32; the CTA-per-item kernel's partials are K^2 / 2 doubles and more,
        // so it gets larger chunks and fewer of them) are above it
        constexpr int MAX_HEAVY = 16384;
        const long long MAX_CHUNKS = K == 32 ? 131072 : (long long)(1.5e9 / (8.0 * block_partial_doubles(K)));
        const int CH = K == 32 ? heavy_chunk_size() : block_heavy_chunk_size();
        const size_t part_doubles = K == 32 ? (size_t)heavy_partial_doubles() : (size_t)block_partial_doubles(K);
        long long thr = std::max<long long>(ctx->heavy_threshold, 16 * (nnz / num_items + 1));
        for (;; thr *= 2) {
            long long cnt = 0, chunks = 0;
            for (int i = 0; i < num_items; ++i) {
                const int64_t len = colptr[i + 1] - colptr[i];
                if (len > thr) { ++cnt; chunks += (len + CH - 1) / CH; }
            }
            if (cnt <= MAX_HEAVY && chunks <= MAX_CHUNKS) break;
        }
        s.heavy_thr = (int)std::min<long long>(thr, 0x7fffffff);
        std::vector<int> items, first(1, 0);
        std::vector<int64_t> p0, p1;
        for (int i = 0; i < num_items; ++i) {
            if (colptr[i + 1] - colptr[i] <= s.heavy_thr) continue;
            items.push_back(i);
            for (int64_t q = colptr[i]; q < colptr[i + 1]; q += CH) { p0.push_back(q); p1.push_back(std::min<int64_t>(q + CH, colptr[i + 1])); }
            first.push_back((int)p0.size());
        }
        if (!items.empty()) {
            s.n_heavy = (int)items.size();
            s.h_heavy_item = items; s.h_heavy_first = first;
            CU(cudaMalloc(&s.hv_item, sizeof(int) * items.size()));
            CU(cudaMalloc(&s.hv_first, sizeof(int) * first.size()));
            CU(cudaMalloc(&s.hv_p0, sizeof(int64_t) * p0.size()));
            CU(cudaMalloc(&s.hv_p1, sizeof(int64_t) * p1.size()));
            CU(cudaMalloc(&s.hv_partials, sizeof(double) * p0.size() * part_doubles));
            CU(cudaMemcpy(s.hv_item, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(s.hv_first, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(s.hv_p0, p0.data(), sizeof(int64_t) * p0.size(), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(s.hv_p1, p1.data(), sizeof(int64_t) * p1.size(), cudaMemcpyHostToDevice));
        }
    }
    s.loaded = true;
    return BPMF_GPU_OK;
}

int bpmf_gpu_load_test(bpmf_gpu_ctx *ctx, int side, const int64_t *colptr, const int32_t *rowidx, const double *val)
{
    if (!ctx || !side_ok(side) || !colptr) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "load_side must come before load_test");
    const int64_t nnz = colptr[s.num];
    if (colptr[0] != 0 || nnz < 0 || (nnz > 0 && (!rowidx || !val))) return fail(ctx, BPMF_GPU_EINVAL, "bad test CSC arrays");
    std::vector<int32_t> col((size_t)nnz);
    for (int i = 0; i < s.num; ++i) {
        if (colptr[i + 1] < colptr[i]) return fail(ctx, BPMF_GPU_EINVAL, "test colptr not monotone");
        for (int64_t p = colptr[i]; p < colptr[i + 1]; ++p) {
            if (rowidx[p] < 0 || rowidx[p] >= s.num_other) return fail(ctx, BPMF_GPU_EINVAL, "test row index out of range");
            col[(size_t)p] = i;
        }
    }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));       // like load_side: nothing in flight may still read the buffers freed below
    CU(cudaStreamSynchronize(ctx->aux_stream));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    dfree(s.t_colptr); dfree(s.t_rowidx); dfree(s.t_col); dfree(s.t_val); dfree(s.pavg); dfree(s.pm2); dfree(s.pred_partials);
    s.nnz_test = nnz;
    const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
    CU(cudaMalloc(&s.t_colptr, sizeof(int64_t) * ((size_t)s.num + 1)));
    CU(cudaMalloc(&s.t_rowidx, sizeof(int32_t) * n1));
    CU(cudaMalloc(&s.t_col, sizeof(int32_t) * n1));
    CU(cudaMalloc(&s.t_val, sizeof(double) * n1));
    CU(cudaMemcpy(s.t_colptr, colptr, sizeof(int64_t) * ((size_t)s.num + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        CU(cudaMemcpy(s.t_rowidx, rowidx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s.t_col, col.data(), sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s.t_val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    return finish_test(ctx, s);
}

// Pm2 = Pavg = T (sample.cpp:123) and the reduction scratch of predict; the test matrix itself is on the device
static int finish_test(bpmf_gpu_ctx *ctx, SideDev &s)
{
    const size_t n1 = (size_t)(s.nnz_test > 0 ? s.nnz_test : 1);
    CU(cudaMalloc(&s.pavg, sizeof(double) * n1));
    CU(cudaMalloc(&s.pm2, sizeof(double) * n1));
    if (s.nnz_test) {
        CU(cudaMemcpy(s.pavg, s.t_val, sizeof(double) * (size_t)s.nnz_test, cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(s.pm2, s.t_val, sizeof(double) * (size_t)s.nnz_test, cudaMemcpyDeviceToDevice));
    }
    s.pred_blocks = ctx->sm_count * 8;
    CU(cudaMalloc(&s.pred_partials, sizeof(double) * (2 * (size_t)s.pred_blocks + 2)));
    return BPMF_GPU_OK;
}

// ---- N4: both factors' matrices built on the device from one coordinate list (build_kernels.cu) ----------------------
namespace {
struct DevCoo {
    int32_t *row = nullptr, *col = nullptr;
    double *val = nullptr;
    ~DevCoo() { cudaFree(row); cudaFree(col); cudaFree(val); }
};
int upload_coo(bpmf_gpu_ctx *ctx, DevCoo &d, int64_t nnz, const int32_t *row, const int32_t *col, const double *val)
{
    const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
    CU(cudaMalloc(&d.row, sizeof(int32_t) * n1));
    CU(cudaMalloc(&d.col, sizeof(int32_t) * n1));
    CU(cudaMalloc(&d.val, sizeof(double) * n1));
    if (nnz) {
        CU(cudaMemcpy(d.row, row, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.col, col, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    return BPMF_GPU_OK;
}
}  // namespace

int bpmf_gpu_load_coo(bpmf_gpu_ctx *ctx, int num_rows, int num_cols, int64_t nnz, const int32_t *row, const int32_t *col,
                      const double *val)
{
    if (!ctx || num_rows < 1 || num_cols < 1 || nnz < 1 || !row || !col || !val) return BPMF_GPU_EINVAL;
    if (nnz > 0x7fffffffll) return fail(ctx, BPMF_GPU_EINVAL, "more than 2^31 - 1 entries");   // int storage indices, like Eigen's
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->aux_stream));
    DevCoo d;
    int rc = upload_coo(ctx, d, nnz, row, col, val);
    if (rc) return rc;
    std::vector<double> hval;
    for (int side = 0; side < 2; ++side) {                 // movies: columns of the matrix; users: its rows
        SideDev &s = ctx->side[side];
        free_side(s);
        const bool mv = side == BPMF_GPU_MOVIES;
        const int num = mv ? num_cols : num_rows, num_other = mv ? num_rows : num_cols;
        int64_t n_out = 0;
        bool bad = false;
        const cudaError_t e = build_compressed(ctx, nnz, num, num_other, mv ? d.col : d.row, mv ? d.row : d.col, d.val, &s.colptr,
                                               &s.rowidx, nullptr, &s.val, &n_out, &bad);
        if (e != cudaSuccess || bad) {
            free_side(s);
            return e != cudaSuccess ? fail(ctx, BPMF_GPU_ECUDA, cudaGetErrorString(e)) : fail(ctx, BPMF_GPU_EINVAL, "matrix entry out of range");
        }
        s.num = num; s.num_other = num_other; s.nnz = n_out;
        // mean_rating = M.sum() / M.nonZeros() (sample.cpp:183): the sum runs over the stored values in storage order, per
        // side, so it is taken on the host from the built array (load time; 8 bytes per entry over PCIe)
        hval.resize((size_t)n_out);
        CU(cudaMemcpy(hval.data(), s.val, sizeof(double) * (size_t)n_out, cudaMemcpyDeviceToHost));
        double sum = 0.0;
        for (double x : hval) sum += x;
        s.mean_rating = sum / (double)n_out;
        std::vector<int64_t> hptr((size_t)num + 1);
        CU(cudaMemcpy(hptr.data(), s.colptr, sizeof(int64_t) * hptr.size(), cudaMemcpyDeviceToHost));
        rc = setup_side(ctx, s, hptr.data());
        if (rc) return rc;
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_load_test_coo(bpmf_gpu_ctx *ctx, int64_t nnz, const int32_t *row, const int32_t *col, const double *val)
{
    if (!ctx || nnz < 0 || (nnz > 0 && (!row || !col || !val))) return BPMF_GPU_EINVAL;
    if (!ctx->side[0].loaded || !ctx->side[1].loaded) return fail(ctx, BPMF_GPU_EINVAL, "the train matrix must be loaded first");
    if (nnz > 0x7fffffffll) return fail(ctx, BPMF_GPU_EINVAL, "more than 2^31 - 1 entries");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    DevCoo d;
    int rc = upload_coo(ctx, d, nnz, row, col, val);
    if (rc) return rc;
    for (int side = 0; side < 2; ++side) {
        SideDev &s = ctx->side[side];
        dfree(s.t_colptr); dfree(s.t_rowidx); dfree(s.t_col); dfree(s.t_val); dfree(s.pavg); dfree(s.pm2); dfree(s.pred_partials);
        const bool mv = side == BPMF_GPU_MOVIES;
        int64_t n_out = 0;
        bool bad = false;
        const cudaError_t e = build_compressed(ctx, nnz, s.num, s.num_other, mv ? d.col : d.row, mv ? d.row : d.col, d.val, &s.t_colptr,
                                               &s.t_rowidx, &s.t_col, &s.t_val, &n_out, &bad);
        if (e != cudaSuccess || bad) {
            dfree(s.t_colptr); dfree(s.t_rowidx); dfree(s.t_col); dfree(s.t_val);
            s.nnz_test = 0;
            return e != cudaSuccess ? fail(ctx, BPMF_GPU_ECUDA, cudaGetErrorString(e)) : fail(ctx, BPMF_GPU_EINVAL, "test entry out of range");
        }
        s.nnz_test = n_out;
        rc = finish_test(ctx, s);
        if (rc) return rc;
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_side(bpmf_gpu_ctx *ctx, int side, int test, int64_t *nnz, double *mean_rating, int64_t *colptr, int32_t *rowidx,
                      double *val)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || (test && !s.t_colptr)) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const int64_t n = test ? s.nnz_test : s.nnz;
    if (nnz) *nnz = n;
    if (mean_rating) *mean_rating = s.mean_rating;
    if (colptr) CU(cudaMemcpy(colptr, test ? s.t_colptr : s.colptr, sizeof(int64_t) * ((size_t)s.num + 1), cudaMemcpyDeviceToHost));
    if (rowidx && n) CU(cudaMemcpy(rowidx, test ? s.t_rowidx : s.rowidx, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    if (val && n) CU(cudaMemcpy(val, test ? s.t_val : s.val, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return BPMF_GPU_OK;
}

int bpmf_gpu_set_range(bpmf_gpu_ctx *ctx, int side, int from, int to)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || from < 0 || to < from || to > s.num) return fail(ctx, BPMF_GPU_EINVAL, "bad range");
    if (from < to && (from < s.slice_from || to > s.slice_to))
        return fail(ctx, BPMF_GPU_EINVAL, "range outside the items whose ratings were loaded (bpmf_gpu_load_side_slice)");
    s.from = from; s.to = to;
    return BPMF_GPU_OK;
}

int bpmf_gpu_bind_items(bpmf_gpu_ctx *ctx, int side, double *dev_items)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    double *target = dev_items ? dev_items : s.items_own;
    if (target != s.items) {
        CU(cudaMemcpyAsync(target, s.items, sizeof(double) * (size_t)ctx->K * s.num, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        s.items = target;
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_set_peers(bpmf_gpu_ctx *ctx, int side, int npeers, double *const *dev_peer_items)
{
    if (!ctx || !side_ok(side) || npeers < 0 || npeers > MAX_PEERS) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    double *tmp[MAX_PEERS] = {nullptr};
    for (int i = 0; i < npeers; ++i) tmp[i] = dev_peer_items[i];
    CU(cudaMemcpyAsync(s.peers_dev, tmp, sizeof(tmp), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < MAX_PEERS; ++i) s.peers_host[i] = tmp[i];
    s.npeers = npeers;
    return BPMF_GPU_OK;
}

int bpmf_gpu_items_device_ptr(bpmf_gpu_ctx *ctx, int side, double **dev_items)
{
    if (!ctx || !side_ok(side) || !dev_items) return BPMF_GPU_EINVAL;
    *dev_items = ctx->side[side].items;
    return BPMF_GPU_OK;
}

int bpmf_gpu_ipc_export(bpmf_gpu_ctx *ctx, int side, unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == BPMF_GPU_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    if (!ctx || !side_ok(side) || !handle) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    if (s.items != s.items_own) return fail(ctx, BPMF_GPU_EINVAL, "externally bound latent storage cannot be exported");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s.items_own));
    memcpy(handle, &h, sizeof h);
    return BPMF_GPU_OK;
}

int bpmf_gpu_ipc_open(bpmf_gpu_ctx *ctx, const unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES], double **dev_items)
{
    if (!ctx || !handle || !dev_items) return BPMF_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_mapped.push_back(p);
    *dev_items = static_cast<double *>(p);
    return BPMF_GPU_OK;
}

int bpmf_gpu_enable_peer_access(bpmf_gpu_ctx *ctx, const bpmf_gpu_ctx *peer)
{
    if (!ctx || !peer) return BPMF_GPU_EINVAL;
    if (ctx->device == peer->device) return BPMF_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    if (!can) return fail(ctx, BPMF_GPU_ECUDA, "devices are not peer-accessible");
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = cudaGetErrorString(e); return BPMF_GPU_ECUDA; }
    (void)cudaGetLastError();
    return BPMF_GPU_OK;
}

int bpmf_gpu_host_alloc(void **host_ptr, uint64_t bytes)
{
    if (!host_ptr) return BPMF_GPU_EINVAL;
    return cudaMallocHost(host_ptr, bytes ? bytes : 1) == cudaSuccess ? BPMF_GPU_OK : BPMF_GPU_ECUDA;
}
int bpmf_gpu_host_free(void *host_ptr) { return cudaFreeHost(host_ptr) == cudaSuccess ? BPMF_GPU_OK : BPMF_GPU_ECUDA; }

int bpmf_gpu_set_items(bpmf_gpu_ctx *ctx, int side, const double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    CU(cudaMemcpyAsync(s.items, host_items, sizeof(double) * (size_t)ctx->K * s.num, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_items(bpmf_gpu_ctx *ctx, int side, double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(host_items, s.items, sizeof(double) * (size_t)ctx->K * s.num, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int bpmf_gpu_set_items_range(bpmf_gpu_ctx *ctx, int side, int from, int to, const double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || from < 0 || to < from || to > s.num) return fail(ctx, BPMF_GPU_EINVAL, "bad range");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    const size_t K = (size_t)ctx->K;
    CU(cudaMemcpyAsync(s.items + K * from, host_items + K * from, sizeof(double) * K * (to - from), cudaMemcpyHostToDevice, ctx->stream));
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_items_range(bpmf_gpu_ctx *ctx, int side, int from, int to, double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || from < 0 || to < from || to > s.num) return fail(ctx, BPMF_GPU_EINVAL, "bad range");
    CU(cudaSetDevice(ctx->device));
    const size_t K = (size_t)ctx->K;
    CU(cudaMemcpyAsync(host_items + K * from, s.items + K * from, sizeof(double) * K * (to - from), cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int bpmf_gpu_push_range(bpmf_gpu_ctx *ctx, int side, int from, int to)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || from < 0 || to < from || to > s.num) return fail(ctx, BPMF_GPU_EINVAL, "bad range");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    const size_t K = (size_t)ctx->K;
    for (int q = 0; q < s.npeers; ++q) {
        double *dst = s.peers_host[q];
        if (!dst || dst == s.items) continue;
        CU(cudaMemcpyAsync(dst + K * from, s.items + K * from, sizeof(double) * K * (to - from), cudaMemcpyDefault, ctx->stream));
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_iter(bpmf_gpu_ctx *ctx, int side, int *iter)
{
    if (!ctx || !side_ok(side) || !iter) return BPMF_GPU_EINVAL;
    *iter = ctx->side[side].iter;
    return BPMF_GPU_OK;
}
int bpmf_gpu_set_iter(bpmf_gpu_ctx *ctx, int side, int iter)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    ctx->side[side].iter = iter;
    return BPMF_GPU_OK;
}

int bpmf_gpu_sample_hyper(bpmf_gpu_ctx *ctx, int side, uint32_t iter, const double *host_sum, const double *host_cov)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    if (s.num < 1) return fail(ctx, BPMF_GPU_EINVAL, "hyper draw needs at least one item");
    CU(cudaSetDevice(ctx->device));
    const int K = ctx->K, KK = K * K;
    if (!host_sum && !host_cov && s.pre_iter == (int)iter) {
        // this iteration's draw was launched right after the previous sweep's reductions (bpmf_gpu_reduce_stats): it has
        // been running on the auxiliary stream under the other side's sweep. Order the stream after it and swap it in.
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_hyper[side], 0));
        std::swap(s.hp, s.hp_next);
        s.pre_iter = -2147483647;
        return BPMF_GPU_OK;
    }
    s.pre_iter = -2147483647;
    const double *d_sum = nullptr, *d_cov = s.cov;
    if (host_sum || host_cov) {
        CU(cudaStreamSynchronize(ctx->stream));  // h_pinned / host_in may still be in use by an earlier call
        if (host_sum) {
            memcpy(ctx->h_pinned, host_sum, sizeof(double) * K);
            CU(cudaMemcpyAsync(ctx->hs[side].host_in, ctx->h_pinned, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
            d_sum = ctx->hs[side].host_in;
        }
        if (host_cov) {
            memcpy(ctx->h_pinned + K, host_cov, sizeof(double) * KK);
            CU(cudaMemcpyAsync(ctx->hs[side].host_in + K, ctx->h_pinned + K, sizeof(double) * KK, cudaMemcpyHostToDevice, ctx->stream));
            d_cov = ctx->hs[side].host_in + K;
        }
    }
    // a stale pre-launched draw of this side may still be running on the auxiliary stream with the same scratch
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_hyper[side], 0));
    CU(launch_hyper(ctx, side, iter, d_sum, d_cov, false));
    CU(cudaEventRecord(ctx->ev_hyper[side], ctx->stream));   // orders a later pre-launch (same scratch) after this draw
    return BPMF_GPU_OK;
}

int bpmf_gpu_set_hyper(bpmf_gpu_ctx *ctx, int side, const double *mu, const double *LambdaF)
{
    if (!ctx || !side_ok(side) || !mu || !LambdaF) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    const int K = ctx->K;
    CU(cudaMemcpyAsync(s.hp.mu, mu, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s.hp.LambdaF, LambdaF, sizeof(double) * K * K, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_hyper(bpmf_gpu_ctx *ctx, int side, double *mu, double *LambdaU, double *LambdaF)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    const int K = ctx->K;
    if (mu) CU(cudaMemcpyAsync(mu, s.hp.mu, sizeof(double) * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (LambdaU) CU(cudaMemcpyAsync(LambdaU, s.hp.LambdaU, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (LambdaF) CU(cudaMemcpyAsync(LambdaF, s.hp.LambdaF, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int bpmf_gpu_sample_items(bpmf_gpu_ctx *ctx, int side, uint32_t iter, double alpha, int kernel_variant)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    const SideDev &o = ctx->side[1 - side];
    if (!s.loaded || !o.loaded) return fail(ctx, BPMF_GPU_EINVAL, "both sides must be loaded");
    if (s.num_other != o.num) return fail(ctx, BPMF_GPU_EINVAL, "sides disagree on dimensions");
    CU(cudaSetDevice(ctx->device));
    const int v = pick_variant(ctx, side, kernel_variant);
    if (s.propLambda && v != BPMF_GPU_KERNEL_EXACT && v != BPMF_GPU_KERNEL_BLOCK && !(v == BPMF_GPU_KERNEL_STREAM && prop_on_stream(ctx, side)))
        return fail(ctx, BPMF_GPU_EINVAL, "a propagated posterior needs the EXACT, AUTO, BLOCK (K = 16 m) or STREAM (K = 32) kernel variant");
    const int slot = (int)(ctx->ev_count % bpmf_gpu_ctx::EV_RING);
    CU(cudaEventRecord(ctx->ev0[slot], ctx->stream));
    if (v == BPMF_GPU_KERNEL_EXACT) CU(launch_items_exact(ctx, side, iter, alpha));
    else if (v == BPMF_GPU_KERNEL_STREAM) {
        if (ctx->K != 32) return fail(ctx, BPMF_GPU_EINVAL, "the stream kernel is built for num_latent == 32");
        CU(launch_items_stream32(ctx, side, iter, alpha));
    } else if (v == BPMF_GPU_KERNEL_BLOCK) {
        if (!block_kernel_supports(ctx->K)) return fail(ctx, BPMF_GPU_EINVAL, "the block kernel is built for num_latent = 16, 48, 64, 80, 96, 112, 128");
        CU(launch_items_block(ctx, side, iter, alpha));
    } else return fail(ctx, BPMF_GPU_EINVAL, "unknown kernel variant");
    CU(cudaEventRecord(ctx->ev1[slot], ctx->stream));
    ctx->ev_count++;
    return BPMF_GPU_OK;
}

int bpmf_gpu_stats_block_items(bpmf_gpu_ctx *ctx, int side, int *items_per_block)
{
    if (!ctx || !side_ok(side) || !items_per_block) return BPMF_GPU_EINVAL;
    if (!ctx->side[side].loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    *items_per_block = stats_block_items(ctx->K, ctx->side[side].num);
    return BPMF_GPU_OK;
}

int bpmf_gpu_stats_block_items_for(int num_latent, int num_items)
{
    if (num_latent < 1 || num_latent > 128 || num_items < 0) return -1;
    return stats_block_items(num_latent, num_items);
}

int bpmf_gpu_stats_device_ptr(bpmf_gpu_ctx *ctx, int side, double **dev_partials)
{
    if (!ctx || !side_ok(side) || !dev_partials) return BPMF_GPU_EINVAL;
    *dev_partials = ctx->side[side].partials;
    return BPMF_GPU_OK;
}

int bpmf_gpu_ipc_export_stats(bpmf_gpu_ctx *ctx, int side, unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES])
{
    if (!ctx || !side_ok(side) || !handle) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s.partials));
    memcpy(handle, &h, sizeof h);
    return BPMF_GPU_OK;
}

int bpmf_gpu_set_stats_peers(bpmf_gpu_ctx *ctx, int side, int npeers, double *const *dev_peer_partials)
{
    if (!ctx || !side_ok(side) || npeers < 0 || npeers > MAX_PEERS || (npeers > 0 && !dev_peer_partials)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    double *tmp[MAX_PEERS] = {nullptr};
    for (int i = 0; i < npeers; ++i) tmp[i] = dev_peer_partials[i];
    CU(cudaMemcpyAsync(s.stat_peers_dev, tmp, sizeof(tmp), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    s.n_stat_peers = npeers;
    s.stat_rank = -1;
    for (int i = 0; i < npeers; ++i)
        if (tmp[i] == s.partials) s.stat_rank = i;
    if (npeers > 0 && s.stat_rank < 0) { s.n_stat_peers = 0; return fail(ctx, BPMF_GPU_EINVAL, "the list must contain this context's own buffer (bpmf_gpu_stats_device_ptr)"); }
    return BPMF_GPU_OK;
}

int bpmf_gpu_peer_barrier(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    if (!ctx->side[side].loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    // once this rank has passed, its peers may overwrite its statistics blocks: the sums that read them must be done
    for (int sd = 0; sd < 2; ++sd) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_sdone[sd], 0));
    CU(launch_peer_barrier(ctx, side, BARRIER_LATENTS, ctx->stream));
    return BPMF_GPU_OK;
}

int bpmf_gpu_reduce_stats_partial(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_sdone[side], 0));   // the previous sums of this side (auxiliary stream) have read the partials
    const cudaError_t e = launch_stats_partial(ctx, side, ctx->stream);
    if (e == cudaErrorInvalidValue && s.n_stat_peers > 0)
        return fail(ctx, BPMF_GPU_EINVAL, "with statistics peers the item range must be aligned to bpmf_gpu_stats_block_items");
    CU(e);
    return BPMF_GPU_OK;
}

int bpmf_gpu_reduce_stats_final(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    // The sums stay on the main stream (two launches of ~5 us). Running them on the auxiliary stream with the hyper draw was
    // measured SLOWER (2 GPUs: 8.16 vs 7.99 ms per step): the persistent item kernel of the other side, enqueued right behind,
    // fills every SM (640 threads x 96 registers leave no room for a second block), so a chain of three small kernels on the
    // auxiliary stream does not get through before that kernel ends and the next sweep of this side waits for the draw. A
    // single hyper kernel launched BEFORE the item kernel does get its SM first; the item kernel's CTA on that SM starts
    // 0.17 ms late and the dynamic claims level it out.
    CU(launch_stats_final(ctx, side, ctx->stream));
    CU(cudaEventRecord(ctx->ev_sdone[side], ctx->stream));
    if (ctx->overlap_hyper && s.num >= 1) {
        // hp.sample of the NEXT iteration needs only this cov (c++/sample.cpp:350): start it now on the auxiliary stream
        CU(cudaEventRecord(ctx->ev_stats[side], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_stats[side], 0));
        CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_hyper[side], 0));
        CU(launch_hyper(ctx, side, (uint32_t)(s.iter + 1), nullptr, s.cov, true));
        CU(cudaEventRecord(ctx->ev_hyper[side], ctx->aux_stream));
        s.pre_iter = s.iter + 1;
    }
    return BPMF_GPU_OK;
}

// The reductions of a sweep on the MAIN stream, behind its item kernel (the host-destination paths, BPMF_STATS_MAIN).
// With statistics peers set this is a COLLECTIVE of the ranks: own blocks to every rank, the cross-GPU barrier (which also
// orders the latent columns the item kernels pushed), the fixed-order sum.
static int reduce_stats_main(bpmf_gpu_ctx *ctx, int side)
{
    int rc = bpmf_gpu_reduce_stats_partial(ctx, side);
    if (rc) return rc;
    if (ctx->side[side].n_stat_peers > 1 && (rc = bpmf_gpu_peer_barrier(ctx, side))) return rc;
    return bpmf_gpu_reduce_stats_final(ctx, side);
}

// The reductions of a sweep (c++/sample.cpp:359-362,379-384, and across GPUs c++/mpi_common.h:44-50). Nothing on the main
// stream needs them before the NEXT sweep of this side, so the whole chain — block partials of the own items (stored into
// every rank's buffer), the cross-GPU barrier of the statistics, the fixed-order sums, cov, and hp.sample of the next
// iteration — runs on the auxiliary stream, under the other side's sweep, on the SMs the item kernels leave free
// (item_sms). The main stream only gets the barrier that orders the pushed latent columns (multi-GPU).
//   Hazards against the peers' stores into this rank's buffers: (1) a peer's next sweep of this side stores statistics
// blocks into `partials` that the sums of THIS sweep read: the peer gets there only after a later latents barrier of the
// OTHER side, and this rank does not arrive at that barrier before the sums are done (ev_sdone of the other side is waited
// for below); (2) the sums wait for every rank's blocks at the statistics barrier. The two kinds of barrier use separate
// arrival words and epochs.
int bpmf_gpu_reduce_stats(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    if (!stats_on_aux(ctx, side) || s.num < 1) return reduce_stats_main(ctx, side);
    CU(cudaSetDevice(ctx->device));
    const bool multi = true;              // (stats_on_aux)
    cudaStream_t aux = ctx->aux_stream;
    CU(cudaEventRecord(ctx->ev_items[side], ctx->stream));
    CU(cudaStreamWaitEvent(aux, ctx->ev_items[side], 0));
    CU(cudaStreamWaitEvent(aux, ctx->ev_hyper[side], 0));        // a draw of this side launched on the main stream (same scratch)
    const cudaError_t e = launch_stats_partial(ctx, side, aux);
    if (e == cudaErrorInvalidValue && s.n_stat_peers > 0)
        return fail(ctx, BPMF_GPU_EINVAL, "with statistics peers the item range must be aligned to bpmf_gpu_stats_block_items");
    CU(e);
    if (multi) CU(launch_peer_barrier(ctx, side, BARRIER_STATS, aux));
    CU(launch_stats_final(ctx, side, aux));
    CU(cudaEventRecord(ctx->ev_sdone[side], aux));
    // hp.sample of the NEXT iteration needs only this cov (c++/sample.cpp:350)
    CU(launch_hyper(ctx, side, (uint32_t)(s.iter + 1), nullptr, s.cov, true));
    CU(cudaEventRecord(ctx->ev_hyper[side], aux));
    s.pre_iter = s.iter + 1;
    if (multi) {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_sdone[1 - side], 0));   // hazard (1)
        CU(launch_peer_barrier(ctx, side, BARRIER_LATENTS, ctx->stream));
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_stats(bpmf_gpu_ctx *ctx, int side, double *sum, double *prod, double *cov, double *norm)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    const int K = ctx->K;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_sdone[side], 0));   // the sums may be in flight on the auxiliary stream
    if (sum) CU(cudaMemcpyAsync(sum, s.sum, sizeof(double) * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (prod) CU(cudaMemcpyAsync(prod, s.prod, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (cov) CU(cudaMemcpyAsync(cov, s.cov, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (norm) CU(cudaMemcpyAsync(norm, s.norm, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int bpmf_gpu_sample(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    s.iter++;                                                             // sample.cpp:344
    int rc = bpmf_gpu_sample_hyper(ctx, side, (uint32_t)s.iter, nullptr, nullptr);  // :349-350 (sum == 0, Q1)
    if (rc) return rc;
    rc = bpmf_gpu_sample_items(ctx, side, (uint32_t)s.iter, alpha, kernel_variant);  // :352-373
    if (rc) return rc;
    if (s.aggrMu && s.iter >= s.aggr_burnin) {                            // :364-368
        rc = bpmf_gpu_aggregate(ctx, side);
        if (rc) return rc;
    }
    return bpmf_gpu_reduce_stats(ctx, side);                              // :379-384
}

int bpmf_gpu_set_prop_posterior(bpmf_gpu_ctx *ctx, int side, const double *host_mu, const double *host_Lambda)
{
    (void)host_mu;   // read and shape-checked by the reference, never used in the draw (c++/sample.cpp:285, quirk Q5)
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    CU(cudaStreamSynchronize(ctx->stream));
    dfree(s.propLambda);
    if (!host_Lambda) return BPMF_GPU_OK;
    const size_t n = sizeof(double) * (size_t)ctx->K * ctx->K * (size_t)(s.num > 0 ? s.num : 1);
    CU(cudaMalloc(&s.propLambda, n));
    CU(cudaMemcpy(s.propLambda, host_Lambda, sizeof(double) * (size_t)ctx->K * ctx->K * s.num, cudaMemcpyHostToDevice));
    return BPMF_GPU_OK;
}

int bpmf_gpu_enable_aggregation(bpmf_gpu_ctx *ctx, int side, int burnin)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded) return fail(ctx, BPMF_GPU_EINVAL, "side not loaded");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    CU(cudaStreamSynchronize(ctx->stream));
    dfree(s.aggrMu); dfree(s.aggrLambda);
    // only the items this context samples: with the items split over G GPUs, K*K*num/G doubles each (c++/bpmf.h:161-176)
    s.aggr_from = s.from; s.aggr_to = s.to;
    const size_t K = (size_t)ctx->K, n = (size_t)(s.to > s.from ? s.to - s.from : 1);
    CU(cudaMalloc(&s.aggrMu, sizeof(double) * K * n));
    CU(cudaMalloc(&s.aggrLambda, sizeof(double) * K * K * n));
    CU(cudaMemset(s.aggrMu, 0, sizeof(double) * K * n));               // sample.cpp:198-199
    CU(cudaMemset(s.aggrLambda, 0, sizeof(double) * K * K * n));
    s.aggr_burnin = burnin;
    return BPMF_GPU_OK;
}

int bpmf_gpu_aggregate(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !s.aggrMu) return fail(ctx, BPMF_GPU_EINVAL, "aggregation is not enabled for this side");
    CU(cudaSetDevice(ctx->device));
    const cudaError_t e = launch_aggregate(ctx, side);
    if (e == cudaErrorInvalidValue) return fail(ctx, BPMF_GPU_EINVAL, "the item range is no longer inside the range aggregation was enabled for");
    CU(e);
    return BPMF_GPU_OK;
}

int bpmf_gpu_finalize_aggregates(bpmf_gpu_ctx *ctx, int side, int nsamples)
{
    // (nsamples <= 1, i.e. fewer iterations than the burn-in, divides by zero exactly as the reference's host code does: inf / nan out)
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !s.aggrMu) return fail(ctx, BPMF_GPU_EINVAL, "aggregation is not enabled for this side");
    CU(cudaSetDevice(ctx->device));
    CU(launch_finalize_aggregates(ctx, side, nsamples));
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_aggregates(bpmf_gpu_ctx *ctx, int side, double *aggrMu, double *aggrLambda)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !s.aggrMu) return fail(ctx, BPMF_GPU_EINVAL, "aggregation is not enabled for this side");
    CU(cudaSetDevice(ctx->device));
    // the host arrays are full size (K x num, K*K x num); only the columns of [aggr_from, aggr_to) are written
    const size_t K = (size_t)ctx->K, n = (size_t)(s.aggr_to - s.aggr_from), off = (size_t)s.aggr_from;
    if (aggrMu && n) CU(cudaMemcpyAsync(aggrMu + K * off, s.aggrMu, sizeof(double) * K * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (aggrLambda && n) CU(cudaMemcpyAsync(aggrLambda + K * K * off, s.aggrLambda, sizeof(double) * K * K * n, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int bpmf_gpu_predict(bpmf_gpu_ctx *ctx, int side, int burnin, double *rmse, double *rmse_avg, int64_t *num_predict)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !ctx->side[1 - side].loaded || !s.t_colptr) return fail(ctx, BPMF_GPU_EINVAL, "test data not loaded");
    CU(cudaSetDevice(ctx->device));
    const int n = (s.iter < burnin) ? 0 : (s.iter - burnin);              // sample.cpp:50
    CU(launch_predict(ctx, side, n));
    double se = 0.0, se_avg = 0.0;
    if (s.nnz_test) {
        CU(cudaMemcpyAsync(ctx->h_pinned, s.pred_partials + 2 * (size_t)s.pred_blocks, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                           ctx->stream));
        const int rc = check_device_error(ctx);
        if (rc) return rc;
        se = ctx->h_pinned[0]; se_avg = ctx->h_pinned[1];
    }
    if (rmse) *rmse = sqrt(se / (double)s.nnz_test);                      // sample.cpp:94-95
    if (rmse_avg) *rmse_avg = sqrt(se_avg / (double)s.nnz_test);
    if (num_predict) *num_predict = s.nnz_test;
    return BPMF_GPU_OK;
}

int bpmf_gpu_get_predictions(bpmf_gpu_ctx *ctx, int side, double *pavg, double *pm2)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !s.t_colptr) return fail(ctx, BPMF_GPU_EINVAL, "test data not loaded");
    CU(cudaSetDevice(ctx->device));
    if (pavg && s.nnz_test) CU(cudaMemcpyAsync(pavg, s.pavg, sizeof(double) * (size_t)s.nnz_test, cudaMemcpyDeviceToHost, ctx->stream));
    if (pm2 && s.nnz_test) CU(cudaMemcpyAsync(pm2, s.pm2, sizeof(double) * (size_t)s.nnz_test, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int64_t bpmf_gpu_launch_count(const bpmf_gpu_ctx *ctx) { return ctx ? ctx->launches : -1; }

int bpmf_gpu_last_items_kernel_ms(bpmf_gpu_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return BPMF_GPU_EINVAL;
    if (!ctx->ev_count) return fail(ctx, BPMF_GPU_EINVAL, "no item kernel has run yet");
    CU(cudaSetDevice(ctx->device));
    const int slot = (int)((ctx->ev_count - 1) % bpmf_gpu_ctx::EV_RING);
    CU(cudaEventSynchronize(ctx->ev1[slot]));
    CU(cudaEventElapsedTime(ms, ctx->ev0[slot], ctx->ev1[slot]));
    return BPMF_GPU_OK;
}

int bpmf_gpu_items_kernel_time(bpmf_gpu_ctx *ctx, double *total_ms, int *count)
{
    if (!ctx || !total_ms || !count) return BPMF_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    long long first = ctx->ev_read;
    if (ctx->ev_count - first > bpmf_gpu_ctx::EV_RING) first = ctx->ev_count - bpmf_gpu_ctx::EV_RING;  // older ones were overwritten
    double tot = 0.0;
    int n = 0;
    for (long long i = first; i < ctx->ev_count; ++i) {
        const int slot = (int)(i % bpmf_gpu_ctx::EV_RING);
        float ms = 0.f;
        CU(cudaEventSynchronize(ctx->ev1[slot]));
        CU(cudaEventElapsedTime(&ms, ctx->ev0[slot], ctx->ev1[slot]));
        tot += ms; ++n;
    }
    ctx->ev_read = ctx->ev_count;
    *total_ms = tot; *count = n;
    return BPMF_GPU_OK;
}

// how many parts a host-destination sweep of n items is cut into, and the sum of their weights: parts of geometrically
// decreasing size (ratio 0.6, below the copy : sweep time ratio of a part on PCIe 5), the last one at least MIN_PART items:
// part p's download hides behind part p + 1's sweep and only the last, smallest part's copy is exposed (6 parts: 3.3 %)
static int host_parts(int n, double *wsum_out)
{
    constexpr int PMAX = bpmf_gpu_ctx::HOST_PARTS, MIN_PART = 8192;
    constexpr double RATIO = 0.6;
    int P = 1;
    double wsum = 1.0;
    for (double w = 1.0, ws = 1.0; P < PMAX; ++P, wsum = ws) {
        w *= RATIO; ws += w;
        if (n * (w / ws) < MIN_PART) break;
    }
    *wsum_out = wsum;
    return P;
}

// the items of [from, to) in P parts, each downloaded on the copy stream while the next one is being sampled
static int sample_parts_to_host(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant, double *host_items, int P, double wsum)
{
    SideDev &s = ctx->side[side];
    const size_t K = (size_t)ctx->K;
    const int from = s.from, to = s.to;
    constexpr double RATIO = 0.6;
    double wdone = 0.0, w = 1.0;
    int lo = from, rc = BPMF_GPU_OK;
    // parts end on statistics-block boundaries when peers are set (so that a part is a valid range for every stage)
    for (int part = 0; part < P && !rc; ++part, w *= RATIO) {
        wdone += w;
        const int hi = part == P - 1 ? to : from + (int)((to - from) * (wdone / wsum));
        s.from = lo; s.to = hi;
        rc = bpmf_gpu_sample_items(ctx, side, (uint32_t)s.iter, alpha, kernel_variant);
        if (rc) break;
        cudaError_t e = cudaEventRecord(ctx->ev_part[part], ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_part[part], 0);
        if (e == cudaSuccess && hi > lo)
            e = cudaMemcpyAsync(host_items + K * lo, s.items + K * lo, sizeof(double) * K * (hi - lo), cudaMemcpyDeviceToHost, ctx->copy_stream);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = BPMF_GPU_ECUDA; }
        lo = hi;
    }
    s.from = from; s.to = to;
    return rc;
}

int bpmf_gpu_sample_host(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant, const double *host_other_items,
                         double *host_items)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    SideDev &o = ctx->side[1 - side];
    if (!s.loaded || !o.loaded) return fail(ctx, BPMF_GPU_EINVAL, "both sides must be loaded");
    CU(cudaSetDevice(ctx->device));
    // (only a multi-GPU sweep leaves reductions that READ the latent matrices on the auxiliary stream; on one GPU the stream
    //  holds nothing but the other side's next hyper draw, which this call must not wait for: 0.12 ms per sweep)
    if (stats_on_aux(ctx, 0) || stats_on_aux(ctx, 1)) { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    const size_t K = (size_t)ctx->K;
    if (s.n_stat_peers > 1) {
        // One rank of a multi-GPU run: host memory holds THIS rank's slice of each latent matrix. Its slice of the other side
        // goes up in chunks, each chunk on to the peers while the next is uploaded; barrier; the sweep in parts, each part
        // downloaded while the next is sampled; own statistics blocks to all ranks; barrier; sum; wait for the download.
        int rc = BPMF_GPU_OK;
        if (host_other_items && (rc = bpmf_gpu_upload_push_range(ctx, 1 - side, o.from, o.to, host_other_items))) return rc;
        if ((rc = bpmf_gpu_peer_barrier(ctx, side))) return rc;
        if (!host_items) return bpmf_gpu_sample(ctx, side, alpha, kernel_variant);
        if ((rc = bpmf_gpu_sample_host_begin(ctx, side, alpha, kernel_variant, host_items))) return rc;
        if ((rc = bpmf_gpu_peer_barrier(ctx, side))) return rc;
        return bpmf_gpu_sample_host_end(ctx, side);
    }
    if (host_other_items)   // other.items() lives in host memory in the reference (bpmf.h:193-194)
        CU(cudaMemcpyAsync(o.items, host_other_items, sizeof(double) * K * o.num, cudaMemcpyHostToDevice, ctx->stream));
    const int from = s.from, to = s.to;
    double wsum = 1.0;
    const int P = host_parts(to - from, &wsum);
    if (!host_items || P < 2) {
        const int rc = bpmf_gpu_sample(ctx, side, alpha, kernel_variant);
        if (rc) return rc;
        if (host_items)
            CU(cudaMemcpyAsync(host_items, s.items, sizeof(double) * K * s.num, cudaMemcpyDeviceToHost, ctx->stream));
        return check_device_error(ctx);   // synchronises the stream
    }
    // Large sweep with a host destination: sample the range in P parts and download each part on the copy stream while
    // the next one is being sampled, so only the last part's copy is exposed. Same stages as bpmf_gpu_sample.
    s.iter++;
    int rc = bpmf_gpu_sample_hyper(ctx, side, (uint32_t)s.iter, nullptr, nullptr);
    if (rc) return rc;
    if (from > 0)   // items other contexts sample (multi-GPU ranges): current on the device, copied as they are
        CU(cudaMemcpyAsync(host_items, s.items, sizeof(double) * K * from, cudaMemcpyDeviceToHost, ctx->stream));
    if (to < s.num)
        CU(cudaMemcpyAsync(host_items + K * to, s.items + K * to, sizeof(double) * K * (s.num - to), cudaMemcpyDeviceToHost, ctx->stream));
    rc = sample_parts_to_host(ctx, side, alpha, kernel_variant, host_items, P, wsum);
    if (rc) return rc;
    if (s.aggrMu && s.iter >= s.aggr_burnin) {
        rc = bpmf_gpu_aggregate(ctx, side);
        if (rc) return rc;
    }
    rc = reduce_stats_main(ctx, side);      // (host-destination path: the other side is uploaded next, keep everything on one stream)
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
    return check_device_error(ctx);   // synchronises the stream
}

// ---- the same for one rank of a multi-GPU run: host memory holds THIS rank's slice of each latent matrix --------------
int bpmf_gpu_upload_push_range(bpmf_gpu_ctx *ctx, int side, int from, int to, const double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || from < 0 || to < from || to > s.num) return fail(ctx, BPMF_GPU_EINVAL, "bad range");
    CU(cudaSetDevice(ctx->device));
    { const int qrc = quiesce_aux(ctx); if (qrc) return qrc; }
    const size_t K = (size_t)ctx->K;
    // chunks: the upload of chunk c + 1 (copy stream, PCIe) runs while chunk c goes to the peers (main stream, NVLink)
    constexpr int NCH = bpmf_gpu_ctx::HOST_PARTS;
    const int n = to - from, nch = n >= NCH * 4096 ? NCH : 1;
    CU(cudaEventRecord(ctx->ev_copied, ctx->stream));            // the copy stream starts after what the main stream has enqueued so far
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copied, 0));
    for (int c = 0; c < nch; ++c) {
        const int lo = from + (int)((long long)n * c / nch), hi = from + (int)((long long)n * (c + 1) / nch);
        if (hi <= lo) continue;
        CU(cudaMemcpyAsync(s.items + K * lo, host_items + K * lo, sizeof(double) * K * (hi - lo), cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaEventRecord(ctx->ev_part[c], ctx->copy_stream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_part[c], 0));
        for (int q = 0; q < s.npeers; ++q) {
            double *dst = s.peers_host[q];
            if (!dst || dst == s.items) continue;
            CU(cudaMemcpyAsync(dst + K * lo, s.items + K * lo, sizeof(double) * K * (hi - lo), cudaMemcpyDefault, ctx->stream));
        }
    }
    return BPMF_GPU_OK;
}

int bpmf_gpu_sample_host_begin(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant, double *host_items)
{
    if (!ctx || !side_ok(side) || !host_items) return BPMF_GPU_EINVAL;
    SideDev &s = ctx->side[side];
    if (!s.loaded || !ctx->side[1 - side].loaded) return fail(ctx, BPMF_GPU_EINVAL, "both sides must be loaded");
    CU(cudaSetDevice(ctx->device));
    s.iter++;
    int rc = bpmf_gpu_sample_hyper(ctx, side, (uint32_t)s.iter, nullptr, nullptr);
    if (rc) return rc;
    double wsum = 1.0;
    const int P = host_parts(s.to - s.from, &wsum);
    rc = sample_parts_to_host(ctx, side, alpha, kernel_variant, host_items, P, wsum);
    if (rc) return rc;
    if (s.aggrMu && s.iter >= s.aggr_burnin) {
        rc = bpmf_gpu_aggregate(ctx, side);
        if (rc) return rc;
    }
    return bpmf_gpu_reduce_stats_partial(ctx, side);
}

int bpmf_gpu_sample_host_end(bpmf_gpu_ctx *ctx, int side)
{
    if (!ctx || !side_ok(side)) return BPMF_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    const int rc = bpmf_gpu_reduce_stats_final(ctx, side);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
    return check_device_error(ctx);   // synchronises the stream: the fresh slice is in host memory
}

int bpmf_gpu_set_heavy_threshold(bpmf_gpu_ctx *ctx, int64_t num_ratings)
{
    if (!ctx || num_ratings < 1) return BPMF_GPU_EINVAL;
    ctx->heavy_threshold = num_ratings;
    return BPMF_GPU_OK;
}

int bpmf_gpu_debug_set_tuning(bpmf_gpu_ctx *ctx, int stream_cfg)
{
    if (!ctx) return BPMF_GPU_EINVAL;
    // the knob carries two fields: cfg % 1000000 is the kernel configuration, cfg / 1000000 - 1 (when >= 0) the number
    // of items per warp that are claimed in small groups at the end of a sweep
    // cfg / 100000000 (when > 0): guided self-scheduling of the claims, a claim = remaining items / (that many quarters of the
    // resident warp count)
    ctx->stream_guided = stream_cfg / 100000000;
    stream_cfg %= 100000000;
    ctx->stream_cfg = stream_cfg % 1000000;
    ctx->stream_tail = stream_cfg / 1000000 - 1;
    return BPMF_GPU_OK;
}

int bpmf_gpu_debug_set_roles(bpmf_gpu_ctx *ctx, unsigned gram_mask, int stages, int slots, int warps)
{
    if (!ctx || (warps != 16 && warps != 20 && warps != 24)) return BPMF_GPU_EINVAL;
    ctx->v5_gram_mask = gram_mask; ctx->v5_ns = stages; ctx->v5_nslot = slots; ctx->v5_nw = warps;
    return BPMF_GPU_OK;
}

int bpmf_gpu_debug_block_schedule(int num_latent, int block_column, int warp, int *out, int cap_quads)
{
    if (!out || cap_quads < 0) return -1;
    return block_schedule(num_latent, block_column, warp, out, cap_quads);
}

int bpmf_gpu_debug_randn(bpmf_gpu_ctx *ctx, uint32_t c, int n, double *host_out)
{
    if (!ctx || n < 1 || n > 4096 || !host_out) return BPMF_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CU(cudaMalloc(&d, sizeof(double) * n));
    cudaError_t e = launch_debug_randn(ctx, c, n, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, BPMF_GPU_ECUDA, cudaGetErrorString(e));
    return BPMF_GPU_OK;
}

}  // extern "C"
