// cuda_sys.h — the B200 back end of `Sys`, in the slot c++/nocomm.h fills for NO_COMM (selected in c++/bpmf.cpp:19-39):
//
//     #define SYS CUDA_Sys
//     struct CUDA_Sys : Sys { ctor(name, fname, probename); ctor(name, M, P); alloc_and_init(); send_item(); sample(); }
//     Sys::Init / Finalize / sync / Abort
//
// Everything numerical goes through the C ABI of libbpmf_b200.so (include/bpmf_gpu.h); this header only orders the
// calls. "movs" is side 0, "users" side 1 (c++/bpmf.cpp:131-132). One host process drives `ngpus` devices (-g N):
// every device holds a full replica of both latent matrices but only the ratings (and -o aggregates) of its own contiguous
// item range [from,to) of each factor (c++/bpmf.h:161-176), samples that range, and its item kernel stores each fresh
// K-vector straight into every replica over NVLink (bpmf_gpu_set_peers) — that store is what replaces send_item() of the
// MPI / GASPI back ends; the sweep statistics are reduced range by range into every device's buffer and a device-side
// barrier replaces the back ends' MPI_Allreduce / MPI_Barrier (bpmf_gpu_set_stats_peers, bpmf_gpu_peer_barrier).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bpmf_gpu.h"
#include "sys.h"

#define SYS CUDA_Sys

struct CUDA_Sys : public Sys {
    //-- c'tor. With device_build (the default) the train / test files are parsed into entry lists and BOTH factors'
    // compressed matrices are built on device 0 (bpmf_gpu_load_coo: radix sort instead of the host's counting sorts and
    // transposes, c++/io.cpp:282,521 + c++/sample.cpp:133-134; bit-identical arrays); Sys::M / T / Pavg / Pm2, which the
    // rest of the program reads on the host, are downloaded from there. -H builds them on the host as the reference does.
    CUDA_Sys(std::string name, std::string fname, std::string probename);
    CUDA_Sys(std::string name, const SparseMatrixD &M, const SparseMatrixD &P);
    ~CUDA_Sys() override;
    static bool device_build;

    void alloc_and_init() override;
    void send_item(int) override {}     // the item kernel has already stored the column into every replica
    void sample(Sys &in) override;
    void predict(Sys &other, bool all = false) override;
    void finalize_mu_lambda() override;

    // host copy of items() after a sweep: needed for -v dumps and by callers that read items() on the host
    // (the reference's own non-virtual predict would). Off by default: predict runs on the device.
    static bool host_coherent;
    static int ngpus;
    static int kernel_variant;
    void fetch_items();

    // balanced contiguous item ranges over the GPUs: work = fixed + nnz per item (c++/assign.cpp:111 uses 10 + nnz)
    std::vector<int> gpu_dom;

    static std::vector<bpmf_gpu_ctx *> gpus;   // one context per device, shared by both factors

  private:
    int side() const { return name == "users" ? BPMF_GPU_USERS : BPMF_GPU_MOVIES; }
    static void check(bpmf_gpu_ctx *c, int rc, const char *what);
    static void ensure_gpus();
    static bool device_built;          // device 0 already holds both factors' train and test matrices
    void download_matrices();          // Sys::M, T (and the copies of T) of this factor from device 0
};

bool CUDA_Sys::device_build = true;
bool CUDA_Sys::device_built = false;

bool CUDA_Sys::host_coherent = false;
int CUDA_Sys::ngpus = 1;
int CUDA_Sys::kernel_variant = BPMF_GPU_KERNEL_AUTO;
std::vector<bpmf_gpu_ctx *> CUDA_Sys::gpus;

void CUDA_Sys::check(bpmf_gpu_ctx *c, int rc, const char *what)
{
    if (rc == BPMF_GPU_OK) return;
    // "Cholesky failed" is the reference's own message (c++/sample.cpp:308); the others name the failing call
    THROWERROR(std::string(what) + ": " + bpmf_gpu_last_error(c));
}

void cuda_sys_shutdown()
{
    for (bpmf_gpu_ctx *c : CUDA_Sys::gpus) bpmf_gpu_destroy(c);
    CUDA_Sys::gpus.clear();
}

void Sys::Init()
{
    Sys::procid = 0;
    Sys::nprocs = 1;
}

void Sys::Finalize() { cuda_sys_shutdown(); }

void Sys::sync()
{
    for (bpmf_gpu_ctx *c : CUDA_Sys::gpus) {
        const int rc = bpmf_gpu_sync(c);
        if (rc) THROWERROR(std::string("device error: ") + bpmf_gpu_last_error(c));
    }
}

void Sys::Abort(int) { abort(); }

CUDA_Sys::~CUDA_Sys()
{
    if (items_ptr) bpmf_gpu_host_free(items_ptr);
    items_ptr = nullptr;
}

void CUDA_Sys::ensure_gpus()
{
    if (!gpus.empty()) return;
    for (int g = 0; g < ngpus; ++g) {
        bpmf_gpu_ctx *c = nullptr;
        const int rc = bpmf_gpu_create(&c, g, num_latent);
        if (rc) THROWERROR(std::string("bpmf_gpu_create: ") + bpmf_gpu_last_error(nullptr));
        gpus.push_back(c);
    }
    for (int g = 0; g < ngpus; ++g)
        for (int h = 0; h < ngpus; ++h)
            if (g != h) check(gpus[(size_t)g], bpmf_gpu_enable_peer_access(gpus[(size_t)g], gpus[(size_t)h]), "enable_peer_access");
}

namespace {
void split_triplets(const bpmf_host::TripletList &l, std::vector<int32_t> &row, std::vector<int32_t> &col, std::vector<double> &val)
{
    const size_t n = l.t.size();
    row.resize(n); col.resize(n); val.resize(n);
    for (size_t i = 0; i < n; ++i) { row[i] = l.t[i].row; col[i] = l.t[i].col; val[i] = l.t[i].val; }
}
}  // namespace

// "movs": the columns of the train file (c++/bpmf.cpp:131, c++/sample.cpp:112-127)
CUDA_Sys::CUDA_Sys(std::string name_, std::string fname, std::string probename) : Sys(name_)
{
    if (!device_build) {
        bpmf_host::read_matrix(fname, M);
        bpmf_host::read_matrix(probename, T);
        const int64_t rows = std::max(M.rows(), T.rows());
        const int64_t cols = std::max(M.cols(), T.cols());
        M.conservativeResize(rows, cols);
        T.conservativeResize(rows, cols);
        Pm2 = Pavg = Torig = T;
        return;
    }
    bpmf_host::TripletList train, test;
    bpmf_host::read_matrix(fname, train);
    bpmf_host::read_matrix(probename, test);
    const int64_t rows = std::max(train.nrows, test.nrows), cols = std::max(train.ncols, test.ncols);   // both grown to the common shape
    if (rows < 1 || cols < 1 || rows > 0x7fffffff || cols > 0x7fffffff || train.t.empty()) THROWERROR("empty train matrix");
    for (const bpmf_host::Triplet &e : train.t)
        if (e.row < 0 || e.row >= train.nrows || e.col < 0 || e.col >= train.ncols) THROWERROR("matrix entry out of range");
    for (const bpmf_host::Triplet &e : test.t)
        if (e.row < 0 || e.row >= test.nrows || e.col < 0 || e.col >= test.ncols) THROWERROR("matrix entry out of range");
    ensure_gpus();
    bpmf_gpu_ctx *c = gpus[0];
    std::vector<int32_t> r, cc;
    std::vector<double> v;
    split_triplets(train, r, cc, v);
    check(c, bpmf_gpu_load_coo(c, (int)rows, (int)cols, (int64_t)v.size(), r.data(), cc.data(), v.data()), "load_coo");
    split_triplets(test, r, cc, v);
    check(c, bpmf_gpu_load_test_coo(c, (int64_t)v.size(), r.data(), cc.data(), v.data()), "load_test_coo");
    device_built = true;
    M.nrows = rows; M.ncols = cols; T.nrows = rows; T.ncols = cols;
    download_matrices();
    if ((train.refuse_duplicates && M.nonZeros() != train.nonZeros()) || (test.refuse_duplicates && T.nonZeros() != test.nonZeros()))
        THROWERROR("Invalid number of values");                       // c++/io.cpp:284-287
}

// "users": the rows of the train file = the transposes (c++/bpmf.cpp:132, c++/sample.cpp:132-137)
CUDA_Sys::CUDA_Sys(std::string name_, const SparseMatrixD &Mt, const SparseMatrixD &Pt) : Sys(name_)
{
    if (!device_built) {
        M = Mt.transpose();
        Pm2 = Pavg = T = Torig = Pt.transpose();
        if (M.rows() != Pavg.rows() || M.cols() != Pavg.cols()) THROWERROR("train and test matrices differ in shape");
        return;
    }
    M.nrows = Mt.cols(); M.ncols = Mt.rows(); T.nrows = Pt.cols(); T.ncols = Pt.rows();
    download_matrices();
}

void CUDA_Sys::download_matrices()
{
    bpmf_gpu_ctx *c = gpus[0];
    const int s = side();
    struct { SparseMatrixD *m; int test; } parts[] = {{&M, 0}, {&T, 1}};
    for (auto &p : parts) {
        int64_t n = 0;
        check(c, bpmf_gpu_get_side(c, s, p.test, &n, nullptr, nullptr, nullptr, nullptr), "get_side");
        p.m->colptr.assign((size_t)p.m->ncols + 1, 0);
        p.m->rowidx.assign((size_t)n, 0);
        p.m->val.assign((size_t)n, 0.0);
        check(c, bpmf_gpu_get_side(c, s, p.test, nullptr, nullptr, p.m->colptr.data(), p.m->rowidx.data(), p.m->val.data()), "get_side");
    }
    Pm2 = Pavg = Torig = T;
}

void CUDA_Sys::alloc_and_init()
{
    ensure_gpus();
    // pinned, so the sweep's host copies run at PCIe speed (the reference mallocs it: c++/nocomm.h:31)
    void *p = nullptr;
    if (bpmf_gpu_host_alloc(&p, sizeof(double) * (size_t)num_latent * (size_t)num())) THROWERROR("pinned allocation failed");
    items_ptr = static_cast<double *>(p);
    init();

    // contiguous ranges balanced on 64 + nnz per item (c++/assign.cpp:111 uses 10 + nnz), cut on statistics-block boundaries (every GPU reduces whole blocks
    // of the fixed decomposition, bpmf_gpu_reduce_stats_partial)
    gpu_dom.assign((size_t)ngpus + 1, 0);
    {
        const int bi = ngpus > 1 ? bpmf_gpu_stats_block_items_for(num_latent, num()) : 1;
        const double total = 64.0 * num() + (double)nnz();   // an item's tail costs ~64 ratings' worth of gather + Gram (K = 32)
        double acc = 0.0;
        int g = 1;
        for (int i = 0; i < num() && g < ngpus; ++i) {
            acc += 64.0 + nnz(i);
            while (g < ngpus && acc >= total * g / ngpus) {
                const int cut = std::min(num(), ((i + 1 + bi / 2) / bi) * bi);
                gpu_dom[(size_t)g] = std::max(cut, gpu_dom[(size_t)g - 1]);
                ++g;
            }
        }
        for (; g <= ngpus; ++g) gpu_dom[(size_t)g] = num();
        gpu_dom[(size_t)ngpus] = num();
    }

    const int s = side();
    for (int g = 0; g < ngpus; ++g) {
        bpmf_gpu_ctx *c = gpus[(size_t)g];
        const int lo = gpu_dom[(size_t)g], hi = gpu_dom[(size_t)g + 1];
        if (device_built && g == 0) {
            // device 0 built the matrices itself and keeps them (and Pavg = Pm2 = T)
        } else if (ngpus == 1) {
            check(c, bpmf_gpu_load_side(c, s, num(), (int)M.rows(), M.colptr.data(), M.rowidx.data(), M.val.data(), mean_rating), "load_side");
        } else {
            // each GPU holds the ratings of its own items only (c++/bpmf.h:161-176); the latent matrices are full replicas
            std::vector<int64_t> cs((size_t)(hi - lo) + 1);
            const int64_t p0 = M.colptr[(size_t)lo];
            for (int i = lo; i <= hi; ++i) cs[(size_t)(i - lo)] = M.colptr[(size_t)i] - p0;
            check(c, bpmf_gpu_load_side_slice(c, s, num(), (int)M.rows(), lo, hi, cs.data(), M.rowidx.data() + p0, M.val.data() + p0, mean_rating),
                  "load_side_slice");
        }
        if (!(device_built && g == 0)) check(c, bpmf_gpu_load_test(c, s, T.colptr.data(), T.rowidx.data(), T.val.data()), "load_test");
        check(c, bpmf_gpu_set_range(c, s, lo, hi), "set_range");
        if (Sys::odirname.size()) check(c, bpmf_gpu_enable_aggregation(c, s, Sys::burnin), "enable_aggregation");
        if (has_prop_posterior()) check(c, bpmf_gpu_set_prop_posterior(c, s, propMu.data(), propLambda.data()), "set_prop_posterior");
    }
    if (ngpus > 1) {
        std::vector<double *> reps((size_t)ngpus, nullptr);
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_items_device_ptr(gpus[(size_t)g], s, &reps[(size_t)g]), "items_device_ptr");
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_set_peers(gpus[(size_t)g], s, ngpus, reps.data()), "set_peers");
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_stats_device_ptr(gpus[(size_t)g], s, &reps[(size_t)g]), "stats_device_ptr");
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_set_stats_peers(gpus[(size_t)g], s, ngpus, reps.data()), "set_stats_peers");
    }
}

// Sys::sample(Sys &other) (c++/sample.cpp:341-385) on the device(s)
void CUDA_Sys::sample(Sys &in)
{
    (void)in;   // the other factor's latent matrix is already resident in every replica
    iter++;
    const int s = side();
    if (ngpus == 1) {
        bpmf_gpu_ctx *c = gpus[0];
        check(c, bpmf_gpu_set_iter(c, s, iter - 1), "set_iter");
        check(c, bpmf_gpu_sample(c, s, Sys::alpha, kernel_variant), "sample");
    } else {
        // every device draws the same hyper-parameters (cov is replicated), samples its range storing each column into every
        // replica, reduces the statistics blocks of its range into every device's buffer, meets the others at the device-side
        // barrier and sums: all inside bpmf_gpu_sample, enqueued device by device without a host synchronisation
        for (bpmf_gpu_ctx *c : gpus) {
            check(c, bpmf_gpu_set_iter(c, s, iter - 1), "set_iter");
            check(c, bpmf_gpu_sample(c, s, Sys::alpha, kernel_variant), "sample");
        }
    }
    // state the main loop reads after sample(): norm (c++/bpmf.cpp:196), cov + hp for inspection
    bpmf_gpu_ctx *c0 = gpus[0];
    check(c0, bpmf_gpu_get_stats(c0, s, nullptr, nullptr, cov.data(), &norm), "get_stats");   // surfaces "Cholesky failed"
    if (host_coherent || Sys::verbose) fetch_items();
}

void CUDA_Sys::fetch_items()
{
    check(gpus[0], bpmf_gpu_get_items(gpus[0], side(), items_ptr), "get_items");
}

// Sys::predict (c++/sample.cpp:48-96) on device 0 (one process owns every test entry, so `all` changes nothing)
void CUDA_Sys::predict(Sys &other, bool all)
{
    (void)other; (void)all;
    bpmf_gpu_ctx *c = gpus[0];
    int64_t np = 0;
    check(c, bpmf_gpu_predict(c, side(), Sys::burnin, &rmse, &rmse_avg, &np), "predict");
    num_predict = (long)np;
    // the final call (c++/bpmf.cpp:225|242) precedes the Pavg / Pm2 dumps: bring them to the host
    if (all && T.nonZeros()) check(c, bpmf_gpu_get_predictions(c, side(), Pavg.val.data(), Pm2.val.data()), "get_predictions");
}

// Sys::finalize_mu_lambda (c++/bpmf.cpp:281-295) on the device(s): every GPU turns the aggregates of its own items into the
// posterior mean and precision (batched K x K inversions, bpmf_gpu_finalize_aggregates) and writes its columns of the host
// matrices
void CUDA_Sys::finalize_mu_lambda()
{
    const int s = side();
    const int nsamples = Sys::nsims - Sys::burnin;
    for (bpmf_gpu_ctx *c : gpus) check(c, bpmf_gpu_finalize_aggregates(c, s, nsamples), "finalize_aggregates");
    for (bpmf_gpu_ctx *c : gpus) check(c, bpmf_gpu_get_aggregates(c, s, aggrMu.data(), aggrLambda.data()), "get_aggregates");
}
