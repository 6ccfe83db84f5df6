// ref_harness.cpp — TEST INFRASTRUCTURE. Drives the REFERENCE's own code for the hot path — c++/sample.cpp and
// c++/mvnormal.cpp, compiled unmodified from /root/reference by oracle/Makefile (target `ref`) against the stand-in
// headers of oracle/shim/ for the two libraries this image lacks (Eigen 3, Random123) — the way the reference's main does
// (c++/bpmf.cpp:131-138 construction, :184-190 sweeps and predictions), behind a small C API, so that
// tests/test_oracle_vs_reference.py can check the restatement in bpmf_oracle.hpp against what the reference's source
// computes: control flow, RNG consumption, the quirks of SURVEY.md §8a, the running averages of predict.
// What this cannot pin is the order of floating-point operations INSIDE Eigen's kernels (the stand-in uses the textbook
// orders, see shim/Eigen/Dense). One binary per K (BPMF_NUMLATENT is a compile-time constant, c++/bpmf.h:53).
// Built without OpenMP: one thread, deterministic reductions.
#include <chrono>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bpmf.h"
#include "io.h"
#include "nocomm.h"   // the NO_COMM back end: SYS = NC_Sys, Sys::Init / Finalize / sync / Abort

// ---- pieces of translation units that are not compiled here -----------------------------------------------------------
// c++/io.cpp (file formats; needs gzstream): the harness hands matrices over in memory instead
void read_matrix(const std::string &name, Eigen::SparseMatrix<double> &) { throw std::runtime_error("ref harness: no file input (" + name + ")"); }
void read_matrix(const std::string &name, Eigen::MatrixXd &) { throw std::runtime_error("ref harness: no file input (" + name + ")"); }
// c++/counters.cpp:160
double tick() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace {

struct Pair {
    SYS *movies = nullptr, *users = nullptr;
    std::string error;
    SYS &side(int s) { return s == 0 ? *movies : *users; }
    SYS &other(int s) { return s == 0 ? *users : *movies; }
};

SparseMatrixD from_coo(long nrows, long ncols, long n, const int *r, const int *c, const double *v)
{
    std::vector<Eigen::Triplet<double>> t;
    t.reserve((size_t)n);
    for (long i = 0; i < n; ++i) t.emplace_back(r[i], c[i], v[i]);
    SparseMatrixD m(nrows, ncols);
    m.setFromTriplets(t.begin(), t.end());
    return m;
}

std::ofstream &devnull()
{
    static std::ofstream f("/dev/null");
    return f;
}

}  // namespace

extern "C" {

int bpmf_ref_num_latent() { return num_latent; }

// rows index users, columns index movies (the layout of the reference's train file). keep_aggr != 0 plays `-o DIR`.
void *bpmf_ref_create(int nrows, int ncols, long nnz, const int *rows, const int *cols, const double *vals, long ntest, const int *trows,
                      const int *tcols, const double *tvals, int burnin, int nsims, double alpha, int keep_aggr)
{
    Sys::Init();                                   // c++/nocomm.h:19-23 (procid = 0, nprocs = 1), as c++/bpmf.cpp:71
    Sys::nsims = nsims; Sys::burnin = burnin; Sys::update_freq = 1; Sys::alpha = alpha;
    Sys::odirname = keep_aggr ? "ref-harness-output" : "";   // this path only tests it for emptiness (sample.cpp:196,364)
    Sys::os = &devnull(); Sys::dbgs = &devnull();
    Pair *p = new Pair;
    try {
        // The file-reading constructor (sample.cpp:112-127) is entered through its matrix twin (sample.cpp:132-137), which
        // TRANSPOSES its arguments: hand it the transposes so that movies.M is the train matrix itself.
        const SparseMatrixD R = from_coo(nrows, ncols, nnz, rows, cols, vals), T = from_coo(nrows, ncols, ntest, trows, tcols, tvals);
        const SparseMatrixD Rt = R.transpose(), Tt = T.transpose();
        p->movies = new SYS("movs", Rt, Tt);
        p->users = new SYS("users", p->movies->M, p->movies->Pavg);       // c++/bpmf.cpp:132
        p->movies->alloc_and_init();                                      // c++/bpmf.cpp:137-138
        p->users->alloc_and_init();
        // Sys::assign (c++/assign.cpp, not compiled: sparse products and permutations) leaves one process with everything
        p->movies->dom = {0, p->movies->num()};
        p->users->dom = {0, p->users->num()};
    } catch (const std::exception &e) {
        p->error = e.what();
    }
    return p;
}

void bpmf_ref_destroy(void *h)
{
    Pair *p = static_cast<Pair *>(h);
    if (!p) return;
    if (p->movies) free(p->movies->items_ptr);
    if (p->users) free(p->users->items_ptr);
    delete p->movies;
    delete p->users;
    delete p;
}

const char *bpmf_ref_error(void *h) { return static_cast<Pair *>(h)->error.c_str(); }

// movies.sample(users) / users.sample(movies) (c++/bpmf.cpp:184-185). 0 = ok, 1 = an exception (bpmf_ref_error).
int bpmf_ref_sample(void *h, int side)
{
    Pair *p = static_cast<Pair *>(h);
    try {
        p->side(side).sample(p->other(side));
    } catch (const std::exception &e) {
        p->error = e.what();
        return 1;
    }
    return 0;
}

int bpmf_ref_predict(void *h, int side, int all)
{
    Pair *p = static_cast<Pair *>(h);
    try {
        p->side(side).predict(p->other(side), all != 0);
    } catch (const std::exception &e) {
        p->error = e.what();
        return 1;
    }
    return 0;
}

int bpmf_ref_num(void *h, int side) { return static_cast<Pair *>(h)->side(side).num(); }
long bpmf_ref_nnz_test(void *h, int side) { return static_cast<Pair *>(h)->side(side).T.nonZeros(); }

void bpmf_ref_get_items(void *h, int side, double *out)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    std::memcpy(out, s.items_ptr, sizeof(double) * (size_t)num_latent * (size_t)s.num());
}
void bpmf_ref_set_items(void *h, int side, const double *in)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    std::memcpy(s.items_ptr, in, sizeof(double) * (size_t)num_latent * (size_t)s.num());
}

// out[0..5] = rmse, rmse_avg, norm, mean_rating, iter, num_predict
void bpmf_ref_get_scalars(void *h, int side, double *out)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    out[0] = s.rmse; out[1] = s.rmse_avg; out[2] = s.norm; out[3] = s.mean_rating; out[4] = s.iter; out[5] = s.num_predict;
}

void bpmf_ref_get_hyper(void *h, int side, double *mu, double *LambdaU, double *LambdaF)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    for (int i = 0; i < num_latent; ++i) mu[i] = s.hp.mu(i);
    for (int j = 0; j < num_latent; ++j)
        for (int i = 0; i < num_latent; ++i) {
            LambdaU[i + j * num_latent] = s.hp.LambdaU(i, j);
            LambdaF[i + j * num_latent] = s.hp.LambdaF(i, j);
        }
}

void bpmf_ref_get_cov(void *h, int side, double *cov)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    for (int j = 0; j < num_latent; ++j)
        for (int i = 0; i < num_latent; ++i) cov[i + j * num_latent] = s.cov(i, j);
}

// Pavg / Pm2 values in the storage order of the side's test matrix
void bpmf_ref_get_predictions(void *h, int side, double *pavg, double *pm2)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    std::memcpy(pavg, s.Pavg.valuePtr(), sizeof(double) * (size_t)s.Pavg.nonZeros());
    std::memcpy(pm2, s.Pm2.valuePtr(), sizeof(double) * (size_t)s.Pm2.nonZeros());
}

// aggrMu (K x num) and aggrLambda (K*K x num), allocated only with keep_aggr (c++/sample.cpp:196-200)
int bpmf_ref_get_aggregates(void *h, int side, double *aggrMu, double *aggrLambda)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    if (s.aggrMu.size() == 0) return 1;
    std::memcpy(aggrMu, s.aggrMu.data(), sizeof(double) * (size_t)s.aggrMu.size());
    std::memcpy(aggrLambda, s.aggrLambda.data(), sizeof(double) * (size_t)s.aggrLambda.size());
    return 0;
}

// the propagated posterior of -m / -l, as Sys::add_prop_posterior would have read it (c++/sample.cpp:157-174)
void bpmf_ref_set_prop(void *h, int side, const double *mu, const double *lambda)
{
    SYS &s = static_cast<Pair *>(h)->side(side);
    s.propMu.resize(num_latent, s.num());
    s.propLambda.resize(num_latent * num_latent, s.num());
    std::memcpy(s.propMu.data(), mu, sizeof(double) * (size_t)s.propMu.size());
    std::memcpy(s.propLambda.data(), lambda, sizeof(double) * (size_t)s.propLambda.size());
}

// rng_set_pos(c); n x randn()  (c++/mvnormal.cpp:34-43)
void bpmf_ref_randn(unsigned c, int n, double *out)
{
    rng_set_pos(c);
    for (int i = 0; i < n; ++i) out[i] = randn();
}

}  // extern "C"
