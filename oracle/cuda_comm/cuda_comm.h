// cuda_comm.h — THE REFERENCE-SIDE BINDING of INTEGRATION.md §1, as a real file: the communication back end a maintainer
// of ExaScience/bpmf would add next to c++/nocomm.h to run the Gibbs sweep on B200s through libbpmf_b200.so. It needs
// nothing from this repository except include/bpmf_gpu.h and the library. Eigen types on this side, plain pointers across
// the C ABI.
//
// oracle/Makefile (target `ref`) builds the reference executable with it — every translation unit of the reference
// unmodified, except that c++/bpmf.cpp gets the one extra branch of its `#if` ladder that INTEGRATION.md names (applied
// on the fly, sed into the compiler's standard input, never stored) — as oracle/_ref/bpmf_ref_cuda_k<K>, against the
// stand-in Eigen / Random123 headers of oracle/shim/. The reference's own main loop, its own host `predict`
// (c++/sample.cpp:48-96, non-virtual there) and its own file formats then run around the CUDA sweep.
//
// `predict`, `bcast`, the -v dumps and `finalize_mu_lambda` of the reference read HOST state (items(), aggrMu,
// aggrLambda): this back end keeps it coherent after every sweep (one device-to-host copy of the fresh latent matrix;
// with -o also of the aggregates). INTEGRATION.md describes the faster variant with `predict` made virtual.
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "bpmf_gpu.h"

#define SYS CUDA_Sys

struct CUDA_Sys : public Sys {
    CUDA_Sys(std::string name, std::string fname, std::string probename) : Sys(name, fname, probename) {}
    CUDA_Sys(std::string name, const SparseMatrixD &M, const SparseMatrixD &P) : Sys(name, M, P) {}

    static bpmf_gpu_ctx *ctx;                       // one context (one GPU) shared by both factors
    int side() const { return name == "users" ? BPMF_GPU_USERS : BPMF_GPU_MOVIES; }
    void chk(int rc)
    {
        if (rc) throw std::runtime_error(bpmf_gpu_last_error(ctx));     // uncaught, like THROWERROR (c++/error.h)
    }

    virtual void alloc_and_init()                   // c++/nocomm.h:29-33
    {
        if (!ctx && bpmf_gpu_create(&ctx, 0, num_latent)) throw std::runtime_error(bpmf_gpu_last_error(nullptr));
        void *p = nullptr;
        if (bpmf_gpu_host_alloc(&p, sizeof(double) * num_latent * num())) throw std::runtime_error("pinned allocation failed");
        items_ptr = static_cast<double *>(p);       // pinned instead of malloc: the copies below run at PCIe speed
        init();                                     // c++/sample.cpp:179-226: mean_rating, zeroed items, report
        // Eigen::SparseMatrix<double> is compressed-column with int indices (c++/bpmf.h:55): widen the outer index
        std::vector<int64_t> colptr(M.outerIndexPtr(), M.outerIndexPtr() + num() + 1);
        chk(bpmf_gpu_load_side(ctx, side(), num(), (int)M.rows(), colptr.data(), M.innerIndexPtr(), M.valuePtr(), mean_rating));
        std::vector<int64_t> tptr(T.outerIndexPtr(), T.outerIndexPtr() + num() + 1);
        chk(bpmf_gpu_load_test(ctx, side(), tptr.data(), T.innerIndexPtr(), T.valuePtr()));
        if (Sys::odirname.size()) chk(bpmf_gpu_enable_aggregation(ctx, side(), Sys::burnin));
        if (has_prop_posterior()) chk(bpmf_gpu_set_prop_posterior(ctx, side(), propMu.data(), propLambda.data()));   // -m / -l
    }

    virtual void send_item(int) {}                  // the kernel has already written the column everywhere

    virtual void sample(Sys &)                      // c++/sample.cpp:341-385
    {
        iter++;
        chk(bpmf_gpu_set_iter(ctx, side(), iter - 1));
        chk(bpmf_gpu_sample(ctx, side(), alpha, BPMF_GPU_KERNEL_AUTO));
        chk(bpmf_gpu_get_stats(ctx, side(), nullptr, nullptr, cov.data(), &norm));   // also surfaces "Cholesky failed"
        chk(bpmf_gpu_get_items(ctx, side(), items_ptr));                             // host items() for predict / -v / bcast
        if (Sys::odirname.size() && iter >= Sys::burnin)                             // host aggregates for finalize_mu_lambda
            chk(bpmf_gpu_get_aggregates(ctx, side(), aggrMu.data(), aggrLambda.data()));
    }
};

bpmf_gpu_ctx *CUDA_Sys::ctx = nullptr;

void Sys::Init()                                    // c++/nocomm.h:19-23
{
    Sys::procid = 0;
    Sys::nprocs = 1;
}
void Sys::Finalize()
{
    bpmf_gpu_destroy(CUDA_Sys::ctx);
    CUDA_Sys::ctx = nullptr;
}
void Sys::sync()
{
    if (CUDA_Sys::ctx) bpmf_gpu_sync(CUDA_Sys::ctx);
}
void Sys::Abort(int) { abort(); }
