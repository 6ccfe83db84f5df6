// cuda_sys.h — the B200 back end of `Sys`, in the slot c++/nocomm.h fills for NO_COMM (selected in c++/bpmf.cpp:19-39):
//
//     #define SYS CUDA_Sys
//     struct CUDA_Sys : Sys { ctor(name, fname, probename); ctor(name, M, P); alloc_and_init(); send_item(); sample(); }
//     Sys::Init / Finalize / sync / Abort
//
// Everything numerical goes through the C ABI of libbpmf_b200.so (include/bpmf_gpu.h); this header only orders the
// calls. "movs" is side 0, "users" side 1 (c++/bpmf.cpp:131-132). One host process drives `ngpus` devices (-g N):
// every device holds a full replica of both latent matrices and the ratings, samples its own contiguous item range
// [from,to) of each factor, and its item kernel stores each fresh K-vector straight into every replica over NVLink
// (bpmf_gpu_set_peers) — that store is what replaces send_item() of the MPI / GASPI back ends.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bpmf_gpu.h"
#include "sys.h"

#define SYS CUDA_Sys

struct CUDA_Sys : public Sys {
    //-- c'tor
    CUDA_Sys(std::string name, std::string fname, std::string probename) : Sys(name, fname, probename) {}
    CUDA_Sys(std::string name, const SparseMatrixD &M, const SparseMatrixD &P) : Sys(name, M, P) {}
    ~CUDA_Sys() override;

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
};

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

void CUDA_Sys::alloc_and_init()
{
    if (gpus.empty()) {
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
    // pinned, so the sweep's host copies run at PCIe speed (the reference mallocs it: c++/nocomm.h:31)
    void *p = nullptr;
    if (bpmf_gpu_host_alloc(&p, sizeof(double) * (size_t)num_latent * (size_t)num())) THROWERROR("pinned allocation failed");
    items_ptr = static_cast<double *>(p);
    init();

    // contiguous ranges balanced on 12 + nnz per item
    gpu_dom.assign((size_t)ngpus + 1, 0);
    {
        const double total = 12.0 * num() + (double)nnz();
        double acc = 0.0;
        int g = 1;
        for (int i = 0; i < num() && g < ngpus; ++i) {
            acc += 12.0 + nnz(i);
            while (g < ngpus && acc >= total * g / ngpus) gpu_dom[(size_t)g++] = i + 1;
        }
        for (; g <= ngpus; ++g) gpu_dom[(size_t)g] = num();
        gpu_dom[(size_t)ngpus] = num();
    }

    const int s = side();
    for (int g = 0; g < ngpus; ++g) {
        bpmf_gpu_ctx *c = gpus[(size_t)g];
        check(c, bpmf_gpu_load_side(c, s, num(), (int)M.rows(), M.colptr.data(), M.rowidx.data(), M.val.data(), mean_rating), "load_side");
        check(c, bpmf_gpu_load_test(c, s, T.colptr.data(), T.rowidx.data(), T.val.data()), "load_test");
        check(c, bpmf_gpu_set_range(c, s, gpu_dom[(size_t)g], gpu_dom[(size_t)g + 1]), "set_range");
        if (Sys::odirname.size()) check(c, bpmf_gpu_enable_aggregation(c, s, Sys::burnin), "enable_aggregation");
        if (has_prop_posterior()) check(c, bpmf_gpu_set_prop_posterior(c, s, propMu.data(), propLambda.data()), "set_prop_posterior");
    }
    if (ngpus > 1) {
        std::vector<double *> reps((size_t)ngpus, nullptr);
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_items_device_ptr(gpus[(size_t)g], s, &reps[(size_t)g]), "items_device_ptr");
        for (int g = 0; g < ngpus; ++g) check(gpus[(size_t)g], bpmf_gpu_set_peers(gpus[(size_t)g], s, ngpus, reps.data()), "set_peers");
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
        // every device draws the same hyper-parameters (cov is replicated), samples its range and pushes the columns
        for (bpmf_gpu_ctx *c : gpus) {
            check(c, bpmf_gpu_set_iter(c, s, iter), "set_iter");
            check(c, bpmf_gpu_sample_hyper(c, s, (uint32_t)iter, nullptr, nullptr), "sample_hyper");
            check(c, bpmf_gpu_sample_items(c, s, (uint32_t)iter, Sys::alpha, kernel_variant), "sample_items");
            if (Sys::odirname.size() && iter >= Sys::burnin) check(c, bpmf_gpu_aggregate(c, s), "aggregate");
        }
        for (bpmf_gpu_ctx *c : gpus) check(c, bpmf_gpu_sync(c), "sync");       // all pushes have landed everywhere
        for (bpmf_gpu_ctx *c : gpus) check(c, bpmf_gpu_reduce_stats(c, s), "reduce_stats");
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

// aggregates come back from the device(s), then the reference's host post-processing (c++/bpmf.cpp:281-295)
void CUDA_Sys::finalize_mu_lambda()
{
    const int s = side();
    const size_t K = (size_t)num_latent;
    if (ngpus == 1) {
        check(gpus[0], bpmf_gpu_get_aggregates(gpus[0], s, aggrMu.data(), aggrLambda.data()), "get_aggregates");
    } else {
        DenseMatrixD mu(num_latent, num()), lam((int64_t)(K * K), num());
        for (int g = 0; g < ngpus; ++g) {
            check(gpus[(size_t)g], bpmf_gpu_get_aggregates(gpus[(size_t)g], s, mu.data(), lam.data()), "get_aggregates");
            const size_t lo = (size_t)gpu_dom[(size_t)g], hi = (size_t)gpu_dom[(size_t)g + 1];
            std::memcpy(aggrMu.data() + lo * K, mu.data() + lo * K, sizeof(double) * K * (hi - lo));
            std::memcpy(aggrLambda.data() + lo * K * K, lam.data() + lo * K * K, sizeof(double) * K * K * (hi - lo));
        }
    }
    Sys::finalize_mu_lambda();
}
