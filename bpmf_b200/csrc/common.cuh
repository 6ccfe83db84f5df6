// common.cuh — context / per-side device state shared by the translation units of libbpmf_b200.so
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace bpmf {

constexpr int MAX_PEERS = 16;
constexpr int BARRIER_LATENTS = 0, BARRIER_STATS = 1;
constexpr int STATS_BLOCKS = 296;  // 2 x 148 SMs; fixed so the reduction order never depends on the GPU count

// error word written by kernels: 0 = ok. code in the high 32 bits, detail (item index) in the low 32.
constexpr unsigned long long ERR_CHOLESKY = 3ull << 32;
constexpr unsigned long long ERR_RNG = 4ull << 32;
constexpr unsigned long long ERR_BARRIER = 6ull << 32;   // a peer did not arrive at the cross-GPU barrier in time

struct HyperDev {         // HyperParams on the device (c++/bpmf.h:78-104)
    double *mu = nullptr;       // K
    double *LambdaU = nullptr;  // K*K upper
    double *LambdaF = nullptr;  // K*K
};

struct SideDev {          // the device mirror of one Sys (c++/bpmf.h:112-239)
    bool loaded = false;
    int num = 0, num_other = 0;
    int64_t nnz = 0, nnz_test = 0;
    int from = 0, to = 0;
    int slice_from = 0, slice_to = 0;     // items whose ratings are resident (bpmf_gpu_load_side_slice); [0, num) after load_side
    int iter = -1;
    double mean_rating = 0.0;
    // train CSC: column = item of this side
    int64_t *colptr = nullptr;
    int32_t *rowidx = nullptr;
    double *val = nullptr;
    // test CSC + expanded column index + running predictions
    int64_t *t_colptr = nullptr;
    int32_t *t_rowidx = nullptr, *t_col = nullptr;
    double *t_val = nullptr, *pavg = nullptr, *pm2 = nullptr;
    // latent matrix
    double *items = nullptr;        // active storage (own or bound)
    double *items_own = nullptr;
    int npeers = 0;
    double **peers_dev = nullptr;   // device array[MAX_PEERS] of replica pointers
    double *peers_host[MAX_PEERS] = {};   // the same pointers, host copy (bpmf_gpu_push_range)
    HyperDev hp;
    HyperDev hp_next;               // written by the pre-launched draw of the next iteration, swapped in when it is consumed
    int pre_iter = -2147483647;     // iteration whose hyper-parameters are (being) drawn into hp_next; INT_MIN + 1 = none
    // reductions
    double *sum = nullptr, *prod = nullptr, *cov = nullptr, *norm = nullptr;
    double *partials = nullptr;     // STATS_BLOCKS x (K*K + K + 1)
    // multi-GPU: every rank reduces only the statistics blocks of ITS item range and stores them into every rank's
    // `partials` (peer-mapped pointers, bpmf_gpu_set_stats_peers); 0 = reduce the full replica locally
    int n_stat_peers = 0;
    double **stat_peers_dev = nullptr;    // device array[MAX_PEERS]
    // cross-GPU barrier (peer_barrier_kernel): 2 x MAX_PEERS arrival words behind the partials of every rank's buffer, one
    // set per KIND of barrier (BARRIER_LATENTS on the main stream, BARRIER_STATS on the auxiliary stream: they interleave
    // differently on different ranks); this context is rank `stat_rank` of the peer list; an epoch counts the barriers of
    // its kind on this side
    int stat_rank = -1;
    unsigned long long barrier_epoch[2] = {0, 0};
    double *pred_partials = nullptr;
    int pred_blocks = 0;
    // propagated posterior (-m / -l): per-item prior precision K*K x num, nullptr = none (bpmf_gpu_set_prop_posterior)
    double *propLambda = nullptr;
    // posterior aggregation (-o): K x num and K*K x num, allocated by bpmf_gpu_enable_aggregation
    double *aggrMu = nullptr, *aggrLambda = nullptr;   // items [aggr_from, aggr_to) only: the range at _enable_aggregation time
    int aggr_from = 0, aggr_to = 0;
    int aggr_burnin = 0;
    // heavy items (K == 32 stream kernel): items with more than the threshold of ratings, cut into chunks (stream_kernel.cu)
    int n_heavy = 0;
    int heavy_thr = 0x7fffffff;                       // items with more ratings than this are heavy
    std::vector<int> h_heavy_item, h_heavy_first;     // host copies: item index (ascending), first chunk (n_heavy + 1)
    int *hv_item = nullptr, *hv_first = nullptr;
    int64_t *hv_p0 = nullptr, *hv_p1 = nullptr;       // per chunk: rating range
    double *hv_partials = nullptr;                    // per chunk: partial Gram + rhs in the DMMA layout
    // dynamic work counter for the item kernels
    unsigned int *work_counter = nullptr;
    // K == 32 stream kernel: the right-hand-side weights (val - mean_rating) * alpha of the ratings (sample.cpp:255), which do
    // not change from sweep to sweep, computed once per alpha (8 bytes per rating) instead of per staged rating in every sweep
    double *wval = nullptr;
    double wval_alpha = 0.0;
    bool wval_valid = false;
};

struct HyperScratch {     // global scratch of the single-block hyper kernel
    int nblk = 0;               // pre-generated Philox blocks
    uint32_t *words = nullptr;  // 4*nblk
    unsigned char *acc = nullptr;   // 2*nblk attempt flags (even word offsets)
    int *rank = nullptr;        // 2*nblk: # accepted before this attempt within its class
    int *pos_of_rank = nullptr; // 2 x nblk
    int *row_start = nullptr;   // K+1 (rank of first kept normal of Bartlett row i, and of z), class in row_cls
    int *row_cls = nullptr;     // K+1
    double *mats = nullptr;     // 6 x K*K: X/LU, Tc (transposed work), L, au, spare, spare
    double *vecs = nullptr;     // 4 x K
    int *piv = nullptr;         // K
    double *host_in = nullptr;  // K + K*K staging for host-provided sum / cov
};

}  // namespace bpmf

struct bpmf_gpu_ctx {
    int device = 0;
    int K = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bpmf::SideDev side[2];
    bpmf::HyperScratch hs[2];             // per side: the two sides' draws may be in flight at the same time
    cudaStream_t aux_stream = nullptr;    // the next iteration's hyper draw runs here, under the other side's sweep
    cudaEvent_t ev_stats[2] = {}, ev_hyper[2] = {};
    cudaEvent_t ev_sdone[2] = {};         // the statistics of a side (sum / prod / cov / norm) are complete (they may be computed on aux_stream)
    cudaEvent_t ev_items[2] = {};         // the item kernel (and -o aggregation) of a side's sweep is complete: its statistics may start
    bool overlap_hyper = true;
    // Multi-GPU sweeps (statistics peers set): the whole reduction chain of a sweep (block partials of the own items ->
    // cross-GPU barrier -> sums -> cov -> next hyper draw) runs on the auxiliary stream UNDER the other side's sweep. The
    // persistent item kernels fill every SM they are given (one CTA of 640 threads x 96 registers + 200 KB per SM), so they
    // are launched on sm_count - reserve_sms SMs and the chain gets the rest. On one GPU the reductions cover the whole
    // matrix and are as much work as two SMs do in a sweep: they stay on the main stream, on all SMs (measured).
    bool stats_aux = true;
    int reserve_sms = 2;
    cudaStream_t copy_stream = nullptr;   // bpmf_gpu_sample_host: downloads finished item ranges while the rest is sampled
    static constexpr int HOST_PARTS = 6;   // most parts a host-destination sweep is cut into (capi.cu: bpmf_gpu_sample_host)
    cudaEvent_t ev_part[HOST_PARTS] = {}, ev_copied = nullptr;
    unsigned long long *d_err = nullptr;  // device error word [0]; [1] receives it when the host asks (read-and-reset)
    unsigned long long *h_err = nullptr;  // pinned host copy
    double *h_pinned = nullptr;           // small pinned staging (K*K + K + 8 doubles)
    double *d_zero_row = nullptr;         // 128 doubles of zeros (padding rows of the bulk-copy gather)
    // ring of CUDA-event pairs around the item kernels (bench.py reads the per-launch durations from it)
    static constexpr int EV_RING = 128;
    cudaEvent_t ev0[EV_RING] = {}, ev1[EV_RING] = {};
    long long ev_count = 0;      // item-kernel launches timed so far
    long long ev_read = 0;       // launches already returned by bpmf_gpu_items_kernel_time
    long long launches = 0;
    long long heavy_threshold = 4096;    // items with more ratings go through the chunked path (K == 32)
    int stream_cfg = 0;                   // 0 = default; see launch_items_stream32
    int stream_tail = -1;                 // items per warp claimed in small groups at the end of a sweep; -1 = default
    int stream_guided = 0;                // > 0: guided claims, a claim = remaining / (stream_guided / 4 x resident warps) items
    unsigned v5_gram_mask = 0;            // != 0: the warp-role kernel (stream_kernel.cu, v5) with these Gram warps
    int v5_ns = 2, v5_nslot = 6, v5_nw = 20;
    std::string err;
    std::vector<void *> ipc_mapped;       // peer allocations opened with cudaIpcOpenMemHandle
};

// ---- launchers implemented in the kernel translation units ------------------------------------
namespace bpmf {
// SMs the persistent item kernels are launched on (the rest is left to the statistics chain on the auxiliary stream)
inline bool stats_on_aux(const bpmf_gpu_ctx *c, int side) { return c->stats_aux && c->side[side].n_stat_peers > 1; }
inline int item_sms(const bpmf_gpu_ctx *c, int side) { const int n = c->sm_count - (stats_on_aux(c, side) ? c->reserve_sms : 0); return n < 1 ? 1 : n; }
// exact_kernels.cu (compiled with -fmad=false)
cudaError_t launch_hyper(bpmf_gpu_ctx *c, int side, uint32_t iter, const double *d_sum, const double *d_cov, bool ahead);
cudaError_t launch_items_exact(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha);
cudaError_t launch_stats(bpmf_gpu_ctx *c, int side);
cudaError_t launch_stats_partial(bpmf_gpu_ctx *c, int side, cudaStream_t stream);   // per-block partial sums (own blocks only when stat peers are set)
cudaError_t launch_stats_final(bpmf_gpu_ctx *c, int side, cudaStream_t stream);   // fixed-order sum of the STATS_BLOCKS partials, cov
int stats_block_items(int K, int num);                          // items per statistics block (the granularity of ranges)
cudaError_t launch_peer_barrier(bpmf_gpu_ctx *c, int side, int kind, cudaStream_t stream);   // every rank's earlier work on that stream has landed everywhere
cudaError_t launch_predict(bpmf_gpu_ctx *c, int side, int n);
cudaError_t launch_aggregate(bpmf_gpu_ctx *c, int side);
cudaError_t launch_finalize_aggregates(bpmf_gpu_ctx *c, int side, int nsamples);   // c++/bpmf.cpp:281-295, batched
cudaError_t launch_fetch_error(bpmf_gpu_ctx *c);     // d_err[1] = atomicExch(d_err[0], 0)
cudaError_t launch_debug_randn(bpmf_gpu_ctx *c, uint32_t seed, int n, double *d_out);
size_t exact_items_smem_bytes(int K);
// block_kernel.cu
bool block_kernel_supports(int K);
int block_schedule(int K, int kb, int warp, int *out, int cap_quads);
cudaError_t launch_items_block(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha);
int block_partial_doubles(int K);
int block_heavy_chunk_size();
// stream_kernel.cu
cudaError_t launch_items_stream32(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha);
cudaError_t launch_stats_partial32(bpmf_gpu_ctx *c, int side, int b0, int nb, cudaStream_t stream);
int heavy_chunk_size();
int heavy_partial_doubles();
// build_kernels.cu
cudaError_t build_compressed(bpmf_gpu_ctx *c, int64_t n, int num_major, int num_minor, const int32_t *d_major, const int32_t *d_minor,
                             const double *d_val, int64_t **colptr_out, int32_t **idx_out, int32_t **major_out, double **val_out,
                             int64_t *nnz_out, bool *bad_index);
}  // namespace bpmf
