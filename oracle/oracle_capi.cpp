// oracle_capi.cpp — extern "C" surface of the CPU oracle for ctypes. TEST INFRASTRUCTURE ONLY
// (see bpmf_oracle.hpp header: pinned against the reference's own hot-path sources built with stand-in headers, and at the Philox KAT
// and libstdc++ distribution level).
#include "bpmf_oracle.hpp"

#include <cstdio>

using namespace bpmf_oracle;

namespace {
thread_local std::string g_err;
template <typename F>
int guarded(F f)
{
    try { f(); return 0; }
    catch (const CholeskyFailed &e) { g_err = e.what(); return 2; }
    catch (const std::exception &e) { g_err = e.what(); return 1; }
}
}  // namespace

extern "C" {

const char *bpmf_oracle_last_error() { return g_err.c_str(); }

// ---- RNG known-answer helpers ------------------------------------------------------------------
void bpmf_oracle_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

// rng_set_pos(c) then n raw 32-bit words in MicroURNG delivery order
void bpmf_oracle_words(uint32_t c, int n, uint32_t *out)
{
    Rng r; r.set_pos(c);
    for (int i = 0; i < n; ++i) out[i] = r.urng();
}
// rng_set_pos(c) then n x randn()
void bpmf_oracle_randn(uint32_t c, int n, double *out)
{
    Rng r; r.set_pos(c);
    for (int i = 0; i < n; ++i) out[i] = r.randn();
}
// rng_set_pos(c); std::gamma_distribution<>(alpha)(rng); then n x randn()
void bpmf_oracle_gamma_then_randn(uint32_t c, double alpha, int n, double *gamma_out, double *out)
{
    Rng r; r.set_pos(c);
    std::gamma_distribution<> g(alpha);
    *gamma_out = g(r.urng);
    for (int i = 0; i < n; ++i) out[i] = r.randn();
}

// ---- dense helpers (for unit tests of the oracle itself) ---------------------------------------
int bpmf_oracle_chol_lower(double *A, int K) { return chol_lower_inplace(A, K); }
void bpmf_oracle_inverse(const double *X, int K, double *out) { inverse_lu(X, K, out); }

// ---- hyper-parameter draw: rng_set_pos(iter); hp.sample(N, sum, cov) ---------------------------
int bpmf_oracle_hyper(int K, int N, uint32_t iter, const double *sum, const double *cov, double *mu, double *LambdaU,
                      double *LambdaF)
{
    return guarded([&] {
        Rng r; r.set_pos(iter);
        Hyper hp; hp.resize(K);
        hyper_sample(r, hp, N, sum, cov);
        std::memcpy(mu, hp.mu.data(), sizeof(double) * K);
        std::memcpy(LambdaU, hp.LambdaU.data(), sizeof(double) * K * K);
        std::memcpy(LambdaF, hp.LambdaF.data(), sizeof(double) * K * K);
    });
}

// ---- whole model -------------------------------------------------------------------------------
void *bpmf_oracle_create(int K, int nrows, int ncols, int64_t nnz, const int32_t *rows, const int32_t *cols,
                         const double *vals, int trows, int tcols, int64_t nnz_t, const int32_t *trow, const int32_t *tcol,
                         const double *tval, double alpha, int burnin, int nthreads, int keep_aggr, int no_covariance)
{
    Model *m = nullptr;
    int rc = guarded([&] {
        m = new Model(K, nrows, ncols, nnz, rows, cols, vals, nnz_t, trow, tcol, tval, trows, tcols, keep_aggr != 0,
                      no_covariance != 0);
        m->alpha = alpha; m->burnin = burnin; m->nthreads = nthreads;
    });
    return rc == 0 ? m : nullptr;
}
void bpmf_oracle_destroy(void *h) { delete static_cast<Model *>(h); }

static Side &side_of(void *h, int side) { Model *m = static_cast<Model *>(h); return side == 0 ? m->movies : m->users; }
static Side &other_of(void *h, int side) { Model *m = static_cast<Model *>(h); return side == 0 ? m->users : m->movies; }

int bpmf_oracle_num(void *h, int side) { return side_of(h, side).num(); }
int64_t bpmf_oracle_nnz(void *h, int side) { return side_of(h, side).M.nnz(); }
int64_t bpmf_oracle_nnz_test(void *h, int side) { return side_of(h, side).T.nnz(); }
double bpmf_oracle_mean_rating(void *h, int side) { return side_of(h, side).mean_rating; }
int bpmf_oracle_iter(void *h, int side) { return side_of(h, side).iter; }

// one Sys::sample(other) of one side
int bpmf_oracle_sweep(void *h, int side)
{
    Model *m = static_cast<Model *>(h);
    return guarded([&] { side_of(h, side).sweep(other_of(h, side), m->alpha, m->burnin, m->nthreads); });
}
// Only the per-item draws of a sweep for items [from,to), with whatever hp/iter the side currently has
// (used to time the hot loop on a bounded sample). Does not touch cov/norm.
int bpmf_oracle_sample_range(void *h, int side, int from, int to)
{
    Model *m = static_cast<Model *>(h);
    return guarded([&] {
        Side &s = side_of(h, side);
        const Side &o = other_of(h, side);
        const int K = s.K;
        bool failed = false;
#ifdef _OPENMP
#pragma omp parallel num_threads(m->nthreads > 0 ? m->nthreads : omp_get_max_threads())
#endif
        {
            Rng rng;
            std::vector<double> rr(K), MM((size_t)K * K);
#ifdef _OPENMP
#pragma omp for schedule(guided)
#endif
            for (int i = from; i < to; ++i) {
                if (failed) continue;
                try { s.sample_item(rng, i, o, m->alpha, rr.data(), MM.data()); }
                catch (const CholeskyFailed &) { failed = true; continue; }
                std::memcpy(&s.items[(size_t)i * K], rr.data(), sizeof(double) * K);
            }
        }
        if (failed) throw CholeskyFailed();
    });
}
int bpmf_oracle_predict(void *h, int side)
{
    Model *m = static_cast<Model *>(h);
    return guarded([&] { side_of(h, side).predict(other_of(h, side), m->burnin); });
}
int bpmf_oracle_iterate(void *h) { return guarded([&] { static_cast<Model *>(h)->iterate(); }); }
int bpmf_oracle_finish(void *h) { return guarded([&] { static_cast<Model *>(h)->finish(); }); }

void bpmf_oracle_get_items(void *h, int side, double *out)
{
    Side &s = side_of(h, side);
    std::memcpy(out, s.items.data(), sizeof(double) * s.items.size());
}
void bpmf_oracle_set_items(void *h, int side, const double *in)
{
    Side &s = side_of(h, side);
    std::memcpy(s.items.data(), in, sizeof(double) * s.items.size());
}
void bpmf_oracle_set_iter(void *h, int side, int iter) { side_of(h, side).iter = iter; }
void bpmf_oracle_get_hyper(void *h, int side, double *mu, double *LambdaU, double *LambdaF)
{
    Side &s = side_of(h, side);
    std::memcpy(mu, s.hp.mu.data(), sizeof(double) * s.K);
    std::memcpy(LambdaU, s.hp.LambdaU.data(), sizeof(double) * s.K * s.K);
    std::memcpy(LambdaF, s.hp.LambdaF.data(), sizeof(double) * s.K * s.K);
}
void bpmf_oracle_set_hyper(void *h, int side, const double *mu, const double *LambdaF)
{
    Side &s = side_of(h, side);
    std::memcpy(s.hp.mu.data(), mu, sizeof(double) * s.K);
    std::memcpy(s.hp.LambdaF.data(), LambdaF, sizeof(double) * s.K * s.K);
}
// sum/prod = raw reductions of the last sweep; cov, norm = the members
void bpmf_oracle_get_stats(void *h, int side, double *sum, double *prod, double *cov, double *norm)
{
    Side &s = side_of(h, side);
    if (sum && !s.last_sum.empty()) std::memcpy(sum, s.last_sum.data(), sizeof(double) * s.K);
    if (prod && !s.last_prod.empty()) std::memcpy(prod, s.last_prod.data(), sizeof(double) * s.K * s.K);
    if (cov) std::memcpy(cov, s.cov.data(), sizeof(double) * s.K * s.K);
    if (norm) *norm = s.norm;
}
void bpmf_oracle_set_cov(void *h, int side, const double *cov)
{
    Side &s = side_of(h, side);
    std::memcpy(s.cov.data(), cov, sizeof(double) * s.K * s.K);
}
void bpmf_oracle_get_rmse(void *h, int side, double *rmse, double *rmse_avg, int64_t *nump)
{
    Side &s = side_of(h, side);
    *rmse = s.rmse; *rmse_avg = s.rmse_avg; *nump = s.num_predict;
}
void bpmf_oracle_get_pred(void *h, int side, double *pavg, double *pm2)
{
    Side &s = side_of(h, side);
    std::memcpy(pavg, s.Pavg.data(), sizeof(double) * s.Pavg.size());
    std::memcpy(pm2, s.Pm2.data(), sizeof(double) * s.Pm2.size());
}
// CSC of the side's train matrix as the oracle built it (so tests can feed the identical structure to the GPU path)
void bpmf_oracle_get_csc(void *h, int side, int which /*0=train,1=test*/, int64_t *colptr, int32_t *rowidx, double *val)
{
    Side &s = side_of(h, side);
    const Csc &m = which == 0 ? s.M : s.T;
    std::memcpy(colptr, m.colptr.data(), sizeof(int64_t) * m.colptr.size());
    std::memcpy(rowidx, m.rowidx.data(), sizeof(int32_t) * m.rowidx.size());
    std::memcpy(val, m.val.data(), sizeof(double) * m.val.size());
}
void bpmf_oracle_get_aggr(void *h, int side, double *aggrMu, double *aggrLambda)
{
    Side &s = side_of(h, side);
    if (aggrMu) std::memcpy(aggrMu, s.aggrMu.data(), sizeof(double) * s.aggrMu.size());
    if (aggrLambda) std::memcpy(aggrLambda, s.aggrLambda.data(), sizeof(double) * s.aggrLambda.size());
}
// add_prop_posterior (sample.cpp:157-174): mu is K x num, lambda K*K x num, both item-major
void bpmf_oracle_set_prop(void *h, int side, const double *mu, const double *lambda)
{
    Side &s = side_of(h, side);
    s.propMu.assign(mu, mu + (size_t)s.K * s.num());
    s.propLambda.assign(lambda, lambda + (size_t)s.K * s.K * s.num());
}

int bpmf_oracle_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
