/* bpmf_gpu.h — C ABI of libbpmf_b200.so: the B200 (sm_100a) implementation of the BPMF Gibbs sweep.
 *
 * This is the drop-in boundary for the reference's communication-backend plug-in point
 * (a header that defines `SYS` and a `struct X_Sys : Sys`, selected in c++/bpmf.cpp:19-39 the way
 * c++/nocomm.h:6-37 does for NO_COMM). A `CUDA_Sys : Sys` backend calls exactly these entry points
 * from `alloc_and_init()` and from its override of `virtual void Sys::sample(Sys &in)`
 * (c++/bpmf.h:144,216,219); INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a BPMF_GPU_E* code and
 *     never throws. bpmf_gpu_last_error() gives the message of the last failure on that context.
 *   - "side" 0 = movies (columns of the input file), 1 = users (rows), as in c++/bpmf.cpp:131-132.
 *   - all dense K x K matrices are column-major, latent matrices are item-major K-vectors
 *     (item i at ptr + i*K), exactly `Sys::items_ptr` (c++/bpmf.h:193-194).
 *   - host pointers are borrowed for the duration of the call only. Pointers named dev_* are CUDA
 *     device pointers on the context's device.
 *   - one context drives one GPU; all calls of a context must come from one host thread at a time
 *     (the reference enters Sys::sample from the main thread, c++/bpmf.cpp:184-185).
 *   - work is enqueued on the context's stream (bpmf_gpu_set_stream); calls that return host data
 *     synchronise that stream, the others are asynchronous.
 */
#ifndef BPMF_GPU_H
#define BPMF_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPMF_GPU_OK 0
#define BPMF_GPU_EINVAL 1      /* bad argument / call order                                   */
#define BPMF_GPU_ECUDA 2       /* a CUDA runtime call or kernel failed                         */
#define BPMF_GPU_ECHOLESKY 3   /* "Cholesky failed" (c++/sample.cpp:308): pivot <= 0 for an item */
#define BPMF_GPU_ERNG 4        /* hyper-parameter draw ran out of pre-generated Philox blocks  */
#define BPMF_GPU_ENODEVICE 5   /* no usable sm_100 device (there is NO CPU fallback)           */

#define BPMF_GPU_MOVIES 0
#define BPMF_GPU_USERS 1

/* kernel variants of the per-item conditional update */
#define BPMF_GPU_KERNEL_AUTO 0   /* fastest available for this K                                          */
#define BPMF_GPU_KERNEL_EXACT 1  /* any K; reference summation order, no FMA contraction (debug / fallback) */
/* (2 was round 1's first tensor-core kernel; removed: the STREAM kernel superseded it)                     */
#define BPMF_GPU_KERNEL_STREAM 3 /* K == 32: persistent, cp.async-staged gather ring, DMMA Gram + blocked LDL^T in registers */
#define BPMF_GPU_KERNEL_BLOCK 4  /* K = 16 m, K != 32, K <= 128: one CTA per item, DMMA Gram in registers, tail in shared memory */

typedef struct bpmf_gpu_ctx bpmf_gpu_ctx;

/* ---- life cycle: Sys::Init / Sys::Finalize (c++/nocomm.h:19-27) ------------------------------ */
int bpmf_gpu_create(bpmf_gpu_ctx **out, int device, int num_latent);
int bpmf_gpu_destroy(bpmf_gpu_ctx *ctx);
const char *bpmf_gpu_last_error(const bpmf_gpu_ctx *ctx); /* ctx may be NULL: last create() failure */
int bpmf_gpu_num_latent(const bpmf_gpu_ctx *ctx);
/* cudaStream_t to enqueue on (NULL = legacy default stream). */
int bpmf_gpu_set_stream(bpmf_gpu_ctx *ctx, void *cuda_stream);
/* Sys::sync (c++/nocomm.h:35): wait for everything enqueued so far (the internal auxiliary stream included), report
 * deferred kernel errors. */
int bpmf_gpu_sync(bpmf_gpu_ctx *ctx);

/* ---- data: the Sys constructors + Sys::init (c++/sample.cpp:112-137,179-190) ------------------
 * Train matrix of one side in compressed-column form: column i = item i of this side, rowidx =
 * indices into the OTHER side ascending, explicit zeros kept (Eigen::SparseMatrix<double>, int32
 * inner indices, c++/bpmf.h:55). mean_rating is Sys::mean_rating (c++/sample.cpp:183). Items are
 * zeroed, iter = -1, cov = 0, norm = 0, range = [0, num_items). */
int bpmf_gpu_load_side(bpmf_gpu_ctx *ctx, int side, int num_items, int num_other, const int64_t *colptr,
                       const int32_t *rowidx, const double *val, double mean_rating);
/* The same for a context that samples only the items of [from, to) (one of G GPUs): only the ratings of those items are made
 * resident (c++/bpmf.h:161-176: a node holds its own slice). colptr_slice has to - from + 1 entries starting at 0,
 * rowidx_slice / val_slice colptr_slice[to - from] entries. The range is set to [from, to); bpmf_gpu_set_range must stay
 * inside it. Latent matrices stay full replicas. */
int bpmf_gpu_load_side_slice(bpmf_gpu_ctx *ctx, int side, int num_items, int num_other, int from, int to, const int64_t *colptr_slice,
                             const int32_t *rowidx_slice, const double *val_slice, double mean_rating);
/* Skew handling (K == 32): an item with more than num_ratings ratings (and more than 16x the side's average; the bar
 * is doubled until at most 16384 items per side are above it) is cut into chunks whose partial Gram matrices are computed by separate warps and added in a fixed
 * order, instead of being one warp's work (ChEMBL's hottest target has 110 118 ratings; the reference only has OpenMP's
 * schedule(guided), c++/sample.cpp:352). Default 4096. Takes effect at the next bpmf_gpu_load_side. */
int bpmf_gpu_set_heavy_threshold(bpmf_gpu_ctx *ctx, int64_t num_ratings);
/* Test matrix T of one side, same layout; Pavg = Pm2 = T (c++/sample.cpp:123). */
int bpmf_gpu_load_test(bpmf_gpu_ctx *ctx, int side, const int64_t *colptr, const int32_t *rowidx, const double *val);
/* Both factors' matrices built ON THE DEVICE from one coordinate list: replaces Eigen's
 * setFromTriplets (c++/io.cpp:282,521: compressed columns, inner indices ascending, duplicate entries summed in input
 * order) and the transpose that gives the other factor its matrix (c++/sample.cpp:133-134). row[i] in [0, num_rows)
 * indexes users, col[i] in [0, num_cols) movies, like the reference's train file. Afterwards both sides are loaded as
 * by bpmf_gpu_load_side (bit-identical arrays), each with mean_rating = sum of its stored values in storage order /
 * nnz (c++/sample.cpp:183). At most 2^31 - 1 entries (Eigen's int storage index). */
int bpmf_gpu_load_coo(bpmf_gpu_ctx *ctx, int num_rows, int num_cols, int64_t nnz, const int32_t *row, const int32_t *col,
                      const double *val);
/* The test matrix of both sides from one coordinate list (same shape as the train matrix, c++/sample.cpp:119-123). */
int bpmf_gpu_load_test_coo(bpmf_gpu_ctx *ctx, int64_t nnz, const int32_t *row, const int32_t *col, const double *val);
/* Read a side's matrix back (test != 0: its test matrix): what a host needs for Sys::nnz(), the work-balanced ranges
 * (c++/assign.cpp:111) and the report of Sys::init. Every output pointer may be NULL. */
int bpmf_gpu_get_side(bpmf_gpu_ctx *ctx, int side, int test, int64_t *nnz, double *mean_rating, int64_t *colptr, int32_t *rowidx,
                      double *val);
/* Sys::from()/to() (c++/bpmf.h:171-172): the items this context samples; others are left to peers. */
int bpmf_gpu_set_range(bpmf_gpu_ctx *ctx, int side, int from, int to);
/* Use caller-owned device storage (K * num_items doubles) for a side's latent matrix, e.g. a
 * torch tensor that NCCL all-gathers into; the current contents are copied over. NULL = go back
 * to internal storage. */
int bpmf_gpu_bind_items(bpmf_gpu_ctx *ctx, int side, double *dev_items);
/* Multi-GPU push: device pointers (peer-mapped, one per rank, NULL entries skipped) to every
 * replica of this side's latent matrix; the item kernel stores each fresh K-vector into all of
 * them as it is produced (replaces Sys::send_item, c++/bpmf.h:216). npeers = 0 turns it off. */
int bpmf_gpu_set_peers(bpmf_gpu_ctx *ctx, int side, int npeers, double *const *dev_peer_items);
int bpmf_gpu_items_device_ptr(bpmf_gpu_ctx *ctx, int side, double **dev_items);
/* One process per GPU: export this context's OWN latent storage of a side as a 64-byte CUDA IPC handle, and map a
 * peer process's handle into this process (the result goes into bpmf_gpu_set_peers). Handles of externally bound
 * storage (bpmf_gpu_bind_items) cannot be exported. bpmf_gpu_destroy unmaps what _ipc_open mapped. */
#define BPMF_GPU_IPC_HANDLE_BYTES 64
int bpmf_gpu_ipc_export(bpmf_gpu_ctx *ctx, int side, unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES]);
int bpmf_gpu_ipc_open(bpmf_gpu_ctx *ctx, const unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES], double **dev_items);
/* One process, several GPUs: let this context's device read/write `peer`'s device memory directly, so that
 * bpmf_gpu_items_device_ptr(peer, ...) can be passed to bpmf_gpu_set_peers(ctx, ...). */
int bpmf_gpu_enable_peer_access(bpmf_gpu_ctx *ctx, const bpmf_gpu_ctx *peer);
/* pinned (page-locked) host memory for Sys::items_ptr, so the per-sweep copies run at PCIe speed */
int bpmf_gpu_host_alloc(void **host_ptr, uint64_t bytes);
int bpmf_gpu_host_free(void *host_ptr);

int bpmf_gpu_set_items(bpmf_gpu_ctx *ctx, int side, const double *host_items);
int bpmf_gpu_get_items(bpmf_gpu_ctx *ctx, int side, double *host_items); /* synchronises */
/* Multi-GPU hosts keep one slice of a latent matrix per rank in host memory: upload / download only the items of
 * [from,to) (host_items is the base of the full K x num_items matrix; only that range is touched), and copy a range of
 * this context's replica into every peer replica set with bpmf_gpu_set_peers (NVLink, asynchronous on the stream). */
int bpmf_gpu_set_items_range(bpmf_gpu_ctx *ctx, int side, int from, int to, const double *host_items);
int bpmf_gpu_get_items_range(bpmf_gpu_ctx *ctx, int side, int from, int to, double *host_items); /* synchronises */
int bpmf_gpu_push_range(bpmf_gpu_ctx *ctx, int side, int from, int to);
int bpmf_gpu_get_iter(bpmf_gpu_ctx *ctx, int side, int *iter);
int bpmf_gpu_set_iter(bpmf_gpu_ctx *ctx, int side, int iter);

/* ---- the hot path ---------------------------------------------------------------------------
 * bpmf_gpu_sample == Sys::sample(Sys &other) of NO_COMM (c++/sample.cpp:341-385): iter++, seed,
 * hp.sample(num, sum (always 0), cov), every item in [from,to) drawn by the
 * fused kernel, then sum / prod / norm reduced over ALL items and cov updated on the device.
 * Multi-GPU (bpmf_gpu_set_peers + bpmf_gpu_set_stats_peers on every rank, ranges on statistics-block boundaries): the same
 * call on every rank is the whole sweep — the item kernel stores each fresh column into every replica and a device-side
 * barrier follows on the context's stream; every rank reduces the statistics blocks of its range into all ranks' buffers, a
 * second device-side barrier, the fixed-order sum and the next hyper draw run on an internal auxiliary stream under the
 * other side's sweep (bpmf_gpu_get_stats, bpmf_gpu_sync and the next bpmf_gpu_sample of the side wait for them): bit-identical
 * to one GPU. */
int bpmf_gpu_sample(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant);

/* The same call with HOST-resident latent matrices on both sides, which is literally the reference's
 * Sys::sample(Sys &other): other.items() is read from host memory (host_other_items, K * num_other doubles,
 * uploaded first; NULL = the device copy is current) and this side's fresh items() are written back to
 * host_items (K * num_items doubles; NULL = leave them on the device). Pinned host memory makes both copies
 * asynchronous to the host; the call returns after the download has finished. */
int bpmf_gpu_sample_host(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant, const double *host_other_items,
                         double *host_items);

/* The same for one rank of a multi-GPU run, whose host memory holds ITS slice of each latent matrix (host pointers are the
 * bases of full-size matrices; only the slice is touched):
 *   _upload_push_range   the slice [from, to) of `side` goes host -> device in chunks and each chunk on to every peer replica
 *                        (bpmf_gpu_set_peers) while the next one is uploaded; asynchronous;
 *   [the host's cross-rank barrier: every rank's slices have landed everywhere]
 *   _sample_host_begin   iter++, hyper draw, the context's item range sampled in parts, each part downloaded to host_items
 *                        while the next is sampled, aggregation, the rank's blocks of the statistics (to all ranks);
 *   [the host's cross-rank barrier]
 *   _sample_host_end     statistics summed, the download waited for; synchronises. */
int bpmf_gpu_upload_push_range(bpmf_gpu_ctx *ctx, int side, int from, int to, const double *host_items);
int bpmf_gpu_sample_host_begin(bpmf_gpu_ctx *ctx, int side, double alpha, int kernel_variant, double *host_items);
int bpmf_gpu_sample_host_end(bpmf_gpu_ctx *ctx, int side);

/* The stages of bpmf_gpu_sample, individually callable (multi-GPU hosts put the exchange of the
 * fresh columns between _sample_items and _reduce_stats; tests probe each stage). */
/* rng_set_pos(iter); hp.sample(N, sum, cov)  (c++/sample.cpp:349-350, c++/bpmf.h:98-103,
 * c++/mvnormal.cpp:56-135). host_sum may be NULL (= zeros, what the reference always passes);
 * host_cov NULL = the side's device-resident cov from the last _reduce_stats. */
int bpmf_gpu_sample_hyper(bpmf_gpu_ctx *ctx, int side, uint32_t iter, const double *host_sum, const double *host_cov);
/* override the hyper-parameters (tests) / read them back: mu[K], LambdaU[K*K], LambdaF[K*K] */
int bpmf_gpu_set_hyper(bpmf_gpu_ctx *ctx, int side, const double *mu, const double *LambdaF);
int bpmf_gpu_get_hyper(bpmf_gpu_ctx *ctx, int side, double *mu, double *LambdaU, double *LambdaF);
/* Sys::sample(long idx, Sys &in) for idx in [from,to) (c++/sample.cpp:263-336, 248-258) with the
 * side's current hyper-parameters; `iter` is the value the reference's Sys::iter has inside the call. */
int bpmf_gpu_sample_items(bpmf_gpu_ctx *ctx, int side, uint32_t iter, double alpha, int kernel_variant);
/* sums.combine()/prods.combine()/norms.combine() + cov (c++/sample.cpp:359-362,379-384) over all items. With statistics
 * peers set (K = 32) the chain is enqueued on the internal auxiliary stream behind the side's item kernel, and the context's
 * stream only gets the barrier that orders the pushed latent columns (see bpmf_gpu_sample); BPMF_STATS_MAIN=1 in the
 * environment at bpmf_gpu_create keeps everything on the context's stream. */
int bpmf_gpu_reduce_stats(bpmf_gpu_ctx *ctx, int side);
/* The two halves of _reduce_stats. Multi-GPU hosts call _partial right after _sample_items, then their cross-rank barrier,
 * then _final: with statistics peers set (below) _partial reduces only the blocks of this context's item range and stores
 * each block's partial sums into EVERY rank's buffer over NVLink, so nobody reduces the whole replica and no all-reduce of
 * K*K + K + 1 doubles is needed (replaces the three MPI_Allreduce of c++/mpi_common.h:44-50); _final adds the fixed
 * number of block partials in a fixed order, so the chain is bit-identical for any GPU count. */
int bpmf_gpu_reduce_stats_partial(bpmf_gpu_ctx *ctx, int side);
int bpmf_gpu_reduce_stats_final(bpmf_gpu_ctx *ctx, int side);
/* The cross-rank barrier itself, ON THE DEVICE: one tiny kernel that stores this rank's epoch into an arrival word of every
 * peer's statistics buffer over NVLink (after a system-wide fence) and waits for the peers' epochs in its own. Everything
 * the ranks enqueued on their streams before it (latent columns pushed by the item kernels, statistics blocks, range copies)
 * is visible to what they enqueue after it. No host synchronisation, no NCCL: with statistics peers set, bpmf_gpu_sample,
 * bpmf_gpu_reduce_stats and bpmf_gpu_sample_host run the whole multi-GPU protocol by themselves (every rank must make the
 * same calls). A rank that does not arrive within ~20 s is reported as an error instead of a hang. No-op without peers. */
int bpmf_gpu_peer_barrier(bpmf_gpu_ctx *ctx, int side);
/* Items per statistics block of a side: item ranges (bpmf_gpu_set_range) must start and end on multiples of it (or at
 * num_items) once statistics peers are set. */
int bpmf_gpu_stats_block_items(bpmf_gpu_ctx *ctx, int side, int *items_per_block);
int bpmf_gpu_stats_block_items_for(int num_latent, int num_items);   /* the same before anything is loaded; -1 = bad argument */
/* Every rank's buffer of block partials, like bpmf_gpu_set_peers for the latent matrices: device pointer of this context's
 * buffer, its CUDA IPC handle (map it with bpmf_gpu_ipc_open), and the list of all ranks' buffers. npeers = 0 = every
 * context reduces its full replica by itself. */
int bpmf_gpu_stats_device_ptr(bpmf_gpu_ctx *ctx, int side, double **dev_partials);
int bpmf_gpu_ipc_export_stats(bpmf_gpu_ctx *ctx, int side, unsigned char handle[BPMF_GPU_IPC_HANDLE_BYTES]);
int bpmf_gpu_set_stats_peers(bpmf_gpu_ctx *ctx, int side, int npeers, double *const *dev_peer_partials);
/* any of the out pointers may be NULL. sum[K], prod[K*K], cov[K*K], norm scalar. synchronises. */
int bpmf_gpu_get_stats(bpmf_gpu_ctx *ctx, int side, double *sum, double *prod, double *cov, double *norm);

/* ---- Sys::predict (c++/sample.cpp:48-96) on the device, over all test entries ------------------ */
int bpmf_gpu_predict(bpmf_gpu_ctx *ctx, int side, int burnin, double *rmse, double *rmse_avg, int64_t *num_predict);
int bpmf_gpu_get_predictions(bpmf_gpu_ctx *ctx, int side, double *pavg, double *pm2);

/* ---- propagated posterior of -m / -l (c++/sample.cpp:152-174,272-283) -------------------------------------------
 * Per-item prior precisions: Lambda is K*K x num_items (item i's K x K matrix, column-major, at Lambda + i*K*K), what
 * Sys::add_prop_posterior reads into propLambda. mu (K x num_items) is accepted for symmetry with the reference, which
 * reads and checks propMu but draws with the global hp.mu (c++/sample.cpp:285). host_Lambda == NULL removes the prior.
 * Items are then sampled by the PROP instantiations of the K = 32 stream kernel (heavy items included), by the CTA-per-item
 * kernel (K = 16 m) or by the any-K kernel; KERNEL_AUTO picks. */
int bpmf_gpu_set_prop_posterior(bpmf_gpu_ctx *ctx, int side, const double *host_mu, const double *host_Lambda);

/* ---- posterior aggregation of -o (c++/sample.cpp:195-199,364-368; read back for c++/bpmf.cpp:229-239) ----
 * After _enable_aggregation, every bpmf_gpu_sample whose iteration is >= burnin adds r to aggrMu.col(i) and
 * vec(r r^T) to aggrLambda.col(i) for the items of [from,to) (K*num + K*K*num doubles of device memory).
 * bpmf_gpu_aggregate is that stage on its own, for hosts that drive the stages individually. */
int bpmf_gpu_enable_aggregation(bpmf_gpu_ctx *ctx, int side, int burnin);
int bpmf_gpu_aggregate(bpmf_gpu_ctx *ctx, int side);
/* Sys::finalize_mu_lambda (c++/bpmf.cpp:281-295) on the device, in place, one warp per item: aggrLambda.col(i) becomes the
 * inverse of (aggrLambda.col(i) - aggrMu.col(i) aggrMu.col(i)^T / nsamples) / (nsamples - 1), aggrMu.col(i) /= nsamples. */
int bpmf_gpu_finalize_aggregates(bpmf_gpu_ctx *ctx, int side, int nsamples);
/* The host arrays are full size (K x num_items, K*K x num_items); the columns of the range aggregation was enabled for
 * (the context's [from, to) at that time — the only ones the device holds) are written. */
int bpmf_gpu_get_aggregates(bpmf_gpu_ctx *ctx, int side, double *aggrMu, double *aggrLambda); /* synchronises */

/* ---- introspection --------------------------------------------------------------------------- */
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t bpmf_gpu_launch_count(const bpmf_gpu_ctx *ctx);
/* milliseconds the last bpmf_gpu_sample_items kernel took, from CUDA events on the context's stream */
int bpmf_gpu_last_items_kernel_ms(bpmf_gpu_ctx *ctx, float *ms);
/* sum of the CUDA-event durations (ms) of the bpmf_gpu_sample_items kernels launched since the previous call
 * of this function, and how many there were (at most the last 128 are kept). Synchronises on them. */
int bpmf_gpu_items_kernel_time(bpmf_gpu_ctx *ctx, double *total_ms, int *count);
/* kernel tuning knob of the K == 32 stream kernel: "<version><stages><warps>", e.g. 3216 (16 warps), 6220 (TMA 1-D bulk-copy
 * gather), 14220 (TMA tile::gather4 gather), 15220 (rank-one DMMA column steps); + 100000000 x q: guided claims of
 * remaining / (q/4 x resident warps) items (q = 9: the fixed bulk / tail split); 0 = default. All bit-identical. */
int bpmf_gpu_debug_set_tuning(bpmf_gpu_ctx *ctx, int stream_cfg);
/* warp roles of the K == 32 stream kernel (tuning): bit w of gram_mask makes warp w of the CTA (`warps` = 16, 20 or 24 of them)
 * a Gram warp with a ring of `stages` gather stages, the others tail warps; `slots` transit slots between them.
 * gram_mask == 0 = the library's default configuration. */
int bpmf_gpu_debug_set_roles(bpmf_gpu_ctx *ctx, unsigned gram_mask, int stages, int slots, int warps);
/* device RNG probes for known-answer tests: n normals of the stream rng_set_pos(c) */
int bpmf_gpu_debug_randn(bpmf_gpu_ctx *ctx, uint32_t c, int n, double *host_out);
/* Host only, no GPU needed (tests): the trailing-update schedule of the CTA-per-item kernel (K = 16 m, K != 32) — the
 * quads of 8 x 8 tiles warp `warp` updates in block column `block_column` of the blocked LDL^T, eight ints per quad
 * (offsets in doubles of Lu(I0,kb), Lu(I1,kb), Lu(J0,kb), Lu(J1,kb), A(I0,J0), A(I0,J1), A(I1,J0), A(I1,J1); -1 = no
 * tile). Returns the number of quads (writes at most cap_quads of them), -1 for a bad argument. */
int bpmf_gpu_debug_block_schedule(int num_latent, int block_column, int warp, int *out, int cap_quads);

#ifdef __cplusplus
}
#endif
#endif /* BPMF_GPU_H */
