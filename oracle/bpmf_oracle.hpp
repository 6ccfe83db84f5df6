// bpmf_oracle.hpp — CPU ORACLE for the BPMF Gibbs sweep. TEST INFRASTRUCTURE ONLY.
//
// This is a from-scratch, Eigen-free restatement of the reference algorithm
// (ExaScience/bpmf, NO_COMM + Random123 build). It exists so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs can check and time the CUDA path against it.
// NOTHING under bpmf_b200/ may include, link or call this file.
//
// PARITY STATUS: **pinned against the reference's own source code, bit for bit; unpinned only in the order of
// floating-point operations inside Eigen's kernels.** The reference as a whole cannot be built in this environment
// (Eigen 3 and Random123 are absent, no network) and ships no golden vectors. Its hot-path sources, however —
// c++/sample.cpp and c++/mvnormal.cpp — are compiled UNMODIFIED from /root/reference into oracle/_ref/ against minimal
// stand-in headers for those two libraries (oracle/shim/, oracle/Makefile target `ref`, oracle/ref_harness.cpp), and
// tests/test_oracle_vs_reference.py holds this restatement to them EXACTLY (latents, hyper-parameters, cov, norm,
// RMSEs, Pavg / Pm2, aggregates, propagated posterior, Cholesky failure; K = 10, 16, 32, 48, 64, 128). That pins everything the
// reference's text decides: RNG keying and consumption, the quirks, the averaging of predict. Also pinned
// (tests/test_oracle_kat.py):
//   * Philox4x32-10 against the Random123 kat_vectors and the C++26 [rand.predef] philox4x32
//     10000th-value check (1955073260);
//   * the uniform->normal/gamma transforms by calling this container's libstdc++ <random>
//     (the very code the reference links: std::normal_distribution, std::gamma_distribution)
//     on top of a MicroURNG restated from the published Random123 MicroURNG.hpp semantics;
//   * structural invariants the reference states (U-mu == mean of post-burn-in dumps, results
//     independent of thread count).
// What stays unpinned: dense linear algebra (LLT, LU inverse, triangular solves, products) follows the textbook
// algorithms here AND in the stand-in; real Eigen blocks and vectorises them, i.e. adds in another order, which is
// not reproducible without Eigen. Agreement with a build against real Eigen is therefore expected at fp64 round-off
// (amplified by conditioning), not bitwise.
//
// Reference lines followed (all under /root/reference/c++/):
//   sample.cpp:48-96   Sys::predict            -> Side::predict
//   sample.cpp:179-190 Sys::init               -> Side::init
//   sample.cpp:248-258 Sys::computeMuLambda    -> Side::sample_item (accumulate loop)
//   sample.cpp:263-336 Sys::sample(idx, other) -> Side::sample_item
//   sample.cpp:341-385 Sys::sample(other)      -> Side::sweep
//   mvnormal.cpp:18-47 rng / randn / randu     -> MicroURNG, Rng
//   mvnormal.cpp:56-135 MvNormalChol_prec, WishartUnitChol, WishartChol, NormalWishart,
//                       CondNormalWishart      -> hyper_sample
//   bpmf.h:78-104      HyperParams             -> Hyper
//   bpmf.cpp:180-210, 217-253 main loop order  -> Model::iterate / Model::finish
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace bpmf_oracle {

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; constants as in Random123 philox.h)
// ----------------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += W0; k1 += W1; }
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ----------------------------------------------------------------------------------------------
// r123::MicroURNG<r123::Philox4x32> restated (Random123 MicroURNG.hpp): 32-bit results, the block
// number n is OR-ed into the top word of the counter, and the four words of a block are handed
// out LAST FIRST (rdata[--last_elem]).  rng({{c}},{{42}}) => ctr={c,0,0,0}, key={42,0}
// (mvnormal.cpp:22-23,37).
// ----------------------------------------------------------------------------------------------
struct MicroURNG {
    typedef uint32_t result_type;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    uint32_t c0[4] = {0, 0, 0, 0};
    uint32_t key[2] = {42u, 0u};
    uint32_t rdata[4] = {0, 0, 0, 0};
    uint32_t n = 0;
    int last_elem = 0;
    void reset(uint32_t c)
    {
        c0[0] = c; c0[1] = c0[2] = c0[3] = 0;
        key[0] = 42u; key[1] = 0u;
        n = 0; last_elem = 0;
    }
    result_type operator()()
    {
        if (last_elem == 0) {
            uint32_t c[4] = {c0[0], c0[1], c0[2], c0[3] | n};
            philox4x32_10(c, key, rdata);
            ++n;
            last_elem = 4;
        }
        return rdata[--last_elem];
    }
};

// thread-local stream, exactly like `static thread_local RNG rng` (mvnormal.cpp:23)
struct Rng {
    MicroURNG urng;
    void set_pos(uint32_t c) { urng.reset(c); }                                   // mvnormal.cpp:34-39
    double randn() { return std::normal_distribution<>()(urng); }                 // mvnormal.cpp:41-43
    double randu() { return std::uniform_real_distribution<>(0., 1.0)(urng); }    // mvnormal.cpp:45-47
};

// ----------------------------------------------------------------------------------------------
// Small dense kernels, column-major K x K (Eigen's default storage; bpmf.h:56)
// ----------------------------------------------------------------------------------------------
// Lower Cholesky in place, reading only the lower triangle (Eigen LLT<_, Lower>, unblocked
// left-looking form). Returns -1 on success or the failing pivot index (pivot <= 0).
inline int chol_lower_inplace(double *A, int K)
{
    for (int k = 0; k < K; ++k) {
        double x = A[k + k * K];
        for (int j = 0; j < k; ++j) x -= A[k + j * K] * A[k + j * K];
        if (x <= 0.0) return k;
        x = std::sqrt(x);
        A[k + k * K] = x;
        for (int i = k + 1; i < K; ++i) {
            double s = A[i + k * K];
            for (int j = 0; j < k; ++j) s -= A[i + j * K] * A[k + j * K];
            A[i + k * K] = s / x;
        }
    }
    return -1;
}
// L y = b  (forward), L^T x = y (backward); L = lower triangle of A
inline void solve_lower_inplace(const double *A, int K, double *b)
{
    for (int i = 0; i < K; ++i) {
        double s = b[i];
        for (int j = 0; j < i; ++j) s -= A[i + j * K] * b[j];
        b[i] = s / A[i + i * K];
    }
}
// (terms are subtracted in the order the unknowns become available, j = K-1 down to i+1: the
//  column-oriented form. Eigen's own order inside its blocked triangular solver is not reproducible
//  without Eigen; any order is the same algorithm up to round-off.)
inline void solve_lower_transposed_inplace(const double *A, int K, double *b)
{
    for (int i = K - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = K - 1; j > i; --j) s -= A[j + i * K] * b[j];
        b[i] = s / A[i + i * K];
    }
}
// upper-triangular U x = b (back substitution); U stored full column-major
inline void solve_upper_inplace(const double *U, int K, double *b)
{
    for (int i = K - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = K - 1; j > i; --j) s -= U[i + j * K] * b[j];
        b[i] = s / U[i + i * K];
    }
}
// general inverse by LU with partial pivoting (Eigen PartialPivLU::inverse for K > 4): the result
// is NOT symmetrised (mvnormal.cpp:124).
inline void inverse_lu(const double *X, int K, double *Xinv)
{
    std::vector<double> lu(X, X + (size_t)K * K);
    std::vector<int> piv(K);
    for (int k = 0; k < K; ++k) {
        int p = k;
        double best = std::fabs(lu[k + k * K]);
        for (int i = k + 1; i < K; ++i) {
            const double v = std::fabs(lu[i + k * K]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < K; ++j) std::swap(lu[k + j * K], lu[p + j * K]);
        const double d = lu[k + k * K];
        for (int i = k + 1; i < K; ++i) lu[i + k * K] /= d;
        for (int j = k + 1; j < K; ++j) {
            const double u = lu[k + j * K];
            for (int i = k + 1; i < K; ++i) lu[i + j * K] -= lu[i + k * K] * u;
        }
    }
    // solve LU * Xinv = P * I column by column
    for (int c = 0; c < K; ++c) {
        double *b = Xinv + (size_t)c * K;
        for (int i = 0; i < K; ++i) b[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < K; ++k)
            if (piv[k] != k) std::swap(b[k], b[piv[k]]);
        for (int i = 0; i < K; ++i) {           // unit lower
            double s = b[i];
            for (int j = 0; j < i; ++j) s -= lu[i + j * K] * b[j];
            b[i] = s;
        }
        for (int i = K - 1; i >= 0; --i) {      // upper
            double s = b[i];
            for (int j = i + 1; j < K; ++j) s -= lu[i + j * K] * b[j];
            b[i] = s / lu[i + i * K];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// HyperParams (bpmf.h:78-104) + CondNormalWishart chain (mvnormal.cpp:56-135)
// ----------------------------------------------------------------------------------------------
struct Hyper {
    int K = 0;
    std::vector<double> mu, LambdaF, LambdaU, LambdaL;
    void resize(int k)
    {
        K = k;
        mu.assign(K, 0.0);
        LambdaF.assign((size_t)K * K, 0.0);
        LambdaU.assign((size_t)K * K, 0.0);
        LambdaL.assign((size_t)K * K, 0.0);
    }
};

// hp.sample(N, sum, cov) with mu0 = 0, b0 = 2, WI = I, df = K  (bpmf.h:98-103). `Um` = sum / N.
// The stream must already be positioned by the caller (rng_set_pos(iter), sample.cpp:349).
inline void hyper_sample(Rng &rng, Hyper &hp, int N, const double *sum, const double *cov)
{
    const int K = hp.K;
    const double kappa = 2.0;  // b0
    const int nu = K;          // df
    std::vector<double> Um(K), mu_m(K), mu_c(K);
    // mvnormal.cpp:118-122
    for (int i = 0; i < K; ++i) {
        Um[i] = sum[i] / N;
        mu_m[i] = 0.0 - Um[i];
        mu_c[i] = (kappa * 0.0 + N * Um[i]) / (kappa + N);
    }
    const double kappa_c = kappa + N;
    const double kappa_m = (kappa * N) / (kappa + N);
    // X = T + N*S + kappa_m * mu_m mu_m^T   (T = WI = I), mvnormal.cpp:123
    std::vector<double> X((size_t)K * K), Tc((size_t)K * K);
    for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i)
            X[i + j * K] = ((i == j ? 1.0 : 0.0) + N * cov[i + j * K]) + kappa_m * (mu_m[i] * mu_m[j]);
    inverse_lu(X.data(), K, Tc.data());  // mvnormal.cpp:124
    const int nu_c = nu + N;             // mvnormal.cpp:125

    // WishartChol (mvnormal.cpp:75-84): chol = sigma.llt() reads the lower triangle of T_c
    std::vector<double> L(Tc);
    const int fail = chol_lower_inplace(L.data(), K);
    if (fail >= 0) {
        // Eigen's LLT records NumericalIssue and carries on with a partial factor; the reference never
        // checks it here. We refuse instead of emulating garbage.
        throw std::runtime_error("hyper_sample: T_c is not positive definite");
    }
    // WishartUnitChol (mvnormal.cpp:64-73)
    std::vector<double> au((size_t)K * K, 0.0);
    for (int i = 0; i < K; ++i) {
        std::gamma_distribution<> gam(0.5 * (nu_c - i));
        au[i + i * K] = std::sqrt(2.0 * gam(rng.urng));
        for (int t = 0; t < K - i - 1; ++t) (void)rng.randn();        // VectorXd r = nrandn(...), discarded (:70)
        for (int j = i + 1; j < K; ++j) au[i + j * K] = rng.randn();  // (:71)
    }
    // U = au * chol.matrixU()   (matrixU = L^T), mvnormal.cpp:83
    std::vector<double> &U = hp.LambdaU;
    std::fill(U.begin(), U.end(), 0.0);
    for (int j = 0; j < K; ++j)
        for (int i = 0; i <= j; ++i) {
            double s = 0.0;
            for (int k = i; k <= j; ++k) s += au[i + k * K] * L[j + k * K];  // L^T(k,j) = L(j,k)
            U[i + j * K] = s;
        }
    // MvNormalChol_prec (mvnormal.cpp:56-61)
    std::vector<double> r(K);
    for (int i = 0; i < K; ++i) r[i] = rng.randn();
    solve_upper_inplace(U.data(), K, r.data());
    const double sk = std::sqrt(kappa_c);
    for (int i = 0; i < K; ++i) hp.mu[i] = (r[i] / sk) + mu_c[i];
    // LambdaF = U^T U ; LambdaL = U^T   (bpmf.h:101-102)
    for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i) {
            double s = 0.0;
            const int kmax = std::min(i, j);
            for (int k = 0; k <= kmax; ++k) s += U[k + i * K] * U[k + j * K];
            hp.LambdaF[i + j * K] = s;
            hp.LambdaL[i + j * K] = U[j + i * K];
        }
}

// ----------------------------------------------------------------------------------------------
// Compressed sparse columns with int32 indices, inner indices ascending, duplicates summed,
// explicit zeros kept (Eigen setFromTriplets semantics, io.cpp:282,521).
// ----------------------------------------------------------------------------------------------
struct Csc {
    int nrows = 0, ncols = 0;
    std::vector<int64_t> colptr;
    std::vector<int32_t> rowidx;
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)val.size(); }

    static Csc from_coo(int nrows, int ncols, int64_t n, const int32_t *r, const int32_t *c, const double *v)
    {
        Csc m;
        m.nrows = nrows; m.ncols = ncols;
        std::vector<int64_t> cnt(ncols + 1, 0);
        for (int64_t i = 0; i < n; ++i) {
            if (r[i] < 0 || r[i] >= nrows || c[i] < 0 || c[i] >= ncols) throw std::runtime_error("coo index out of range");
            cnt[c[i] + 1]++;
        }
        for (int j = 0; j < ncols; ++j) cnt[j + 1] += cnt[j];
        std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
        std::vector<int32_t> ri(n);
        std::vector<double> va(n);
        for (int64_t i = 0; i < n; ++i) { const int64_t p = pos[c[i]]++; ri[p] = r[i]; va[p] = v[i]; }
        m.colptr.assign(ncols + 1, 0);
        m.rowidx.reserve(n); m.val.reserve(n);
        std::vector<int64_t> order;
        for (int j = 0; j < ncols; ++j) {
            const int64_t b = cnt[j], e = cnt[j + 1];
            order.resize(e - b);
            for (int64_t t = 0; t < e - b; ++t) order[t] = b + t;
            std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return ri[x] < ri[y]; });
            for (int64_t t = 0; t < e - b; ++t) {
                const int64_t p = order[t];
                if (t > 0 && ri[p] == m.rowidx.back() && (int64_t)m.rowidx.size() > m.colptr[j]) m.val.back() += va[p];
                else { m.rowidx.push_back(ri[p]); m.val.push_back(va[p]); }
            }
            m.colptr[j + 1] = (int64_t)m.val.size();
        }
        return m;
    }
    Csc transpose() const
    {
        Csc t;
        t.nrows = ncols; t.ncols = nrows;
        t.colptr.assign(nrows + 1, 0);
        for (int64_t p = 0; p < nnz(); ++p) t.colptr[rowidx[p] + 1]++;
        for (int i = 0; i < nrows; ++i) t.colptr[i + 1] += t.colptr[i];
        t.rowidx.resize(nnz()); t.val.resize(nnz());
        std::vector<int64_t> pos(t.colptr.begin(), t.colptr.end() - 1);
        for (int j = 0; j < ncols; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int64_t q = pos[rowidx[p]]++;
                t.rowidx[q] = j; t.val[q] = val[p];
            }
        return t;
    }
    void resize_dims(int nr, int nc)  // conservativeResize to LARGER dims (sample.cpp:119-122)
    {
        if (nr < nrows || nc < ncols) throw std::runtime_error("resize_dims only grows");
        nrows = nr;
        colptr.resize(nc + 1, colptr.back());
        ncols = nc;
    }
};

struct CholeskyFailed : std::runtime_error {
    CholeskyFailed() : std::runtime_error("Cholesky failed") {}   // sample.cpp:308
};

// ----------------------------------------------------------------------------------------------
// One factor ("movies" or "users"): the oracle's Sys
// ----------------------------------------------------------------------------------------------
struct Side {
    int K = 0;
    std::string name;
    int iter = -1;                  // sample.cpp:113
    Csc M, T;                       // train / test, column = item of this side
    std::vector<double> Pavg, Pm2;  // same structure as T (sample.cpp:123)
    double mean_rating = 0.0;
    std::vector<double> items;      // K x num(), item-major (bpmf.h:193-194)
    std::vector<double> sum, cov;   // member `sum` is never updated (quirk Q1, sample.cpp:379 shadows it)
    double norm = 0.0;
    Hyper hp;
    double rmse = 0, rmse_avg = 0;
    int64_t num_predict = 0;
    bool keep_aggr = false, no_covariance = false;
    std::vector<double> aggrMu, aggrLambda;
    // propagated posterior (-m / -l, sample.cpp:152-174): per-item prior precision K*K x num(); propMu is read and
    // checked by the reference but never used in the draw (quirk Q5: rr = hp_LambdaF * hp.mu uses the GLOBAL mu)
    std::vector<double> propMu, propLambda;
    bool has_prop_posterior() const { return !propMu.empty(); }
    // last sweep's raw reductions (exposed for parity tests of the device reductions)
    std::vector<double> last_sum, last_prod;

    int num() const { return M.ncols; }

    void init(int k)  // sample.cpp:179-201
    {
        K = k;
        double s = 0.0;
        for (double v : M.val) s += v;
        mean_rating = s / (double)M.nnz();
        items.assign((size_t)K * num(), 0.0);
        sum.assign(K, 0.0);
        cov.assign((size_t)K * K, 0.0);
        norm = 0.0;
        hp.resize(K);
        Pavg = T.val; Pm2 = T.val;
        if (keep_aggr) {
            aggrMu.assign((size_t)K * num(), 0.0);
            aggrLambda.assign((size_t)K * K * num(), 0.0);
        }
    }

    // sample.cpp:263-336 (+ :248-258 inlined). Returns the new latent vector in rr[K].
    void sample_item(Rng &rng, long idx, const Side &other, double alpha, double *rr, double *MM) const
    {
        rng.set_pos((uint32_t)((idx + 1) * (long)K * (long)(iter + 1)));  // :266, 64-bit product truncated
        // hp_LambdaF = propLambda.col(idx) with a propagated posterior, else hp.LambdaF (:272-283)
        const double *LF = has_prop_posterior() ? &propLambda[(size_t)idx * K * K] : hp.LambdaF.data();
        // rr = hp_LambdaF * hp.mu  (:285; hp.mu, not hp_mu: quirk Q5)
        for (int i = 0; i < K; ++i) {
            double s = 0.0;
            for (int j = 0; j < K; ++j) s += LF[i + j * K] * hp.mu[j];
            rr[i] = s;
        }
        std::fill(MM, MM + (size_t)K * K, 0.0);
        // computeMuLambda (:251-257): upper triangle += y y^T ; rr += y * ((v - mean) * alpha)
        for (int64_t p = M.colptr[idx]; p < M.colptr[idx + 1]; ++p) {
            const double *y = &other.items[(size_t)M.rowidx[p] * K];
            const double w = (M.val[p] - mean_rating) * alpha;
            for (int b = 0; b < K; ++b) {
                const double yb = y[b];
                for (int a = 0; a <= b; ++a) MM[a + b * K] += y[a] * yb;
            }
            for (int a = 0; a < K; ++a) rr[a] += y[a] * w;
        }
        // mirror upper -> lower, MM = LambdaF + alpha * MM (:297-298)
        for (int b = 0; b < K; ++b)
            for (int a = b + 1; a < K; ++a) MM[a + b * K] = MM[b + a * K];
        for (size_t e = 0; e < (size_t)K * K; ++e) MM[e] = LF[e] + alpha * MM[e];
        if (no_covariance)  // BPMF_NO_COVARIANCE (:300-304)
            for (int b = 0; b < K; ++b)
                for (int a = 0; a < K; ++a)
                    if (a != b) MM[a + b * K] = 0.0;
        if (chol_lower_inplace(MM, K) >= 0) throw CholeskyFailed();  // :306-308
        solve_lower_inplace(MM, K, rr);                              // :321
        for (int i = 0; i < K; ++i) rr[i] += rng.randn();            // :322
        solve_lower_transposed_inplace(MM, K, rr);                   // :323
    }

    // sample.cpp:341-385
    void sweep(const Side &other, double alpha, int burnin, int nthreads)
    {
        iter++;
        {
            Rng rng;
            rng.set_pos((uint32_t)iter);                      // :349
            hyper_sample(rng, hp, num(), sum.data(), cov.data());  // :350 (member sum == 0, Q1)
        }
        const int N = num();
        std::vector<double> s(K, 0.0), prod((size_t)K * K, 0.0);
        double nrm = 0.0;
        bool failed = false;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
        {
            Rng rng;
            std::vector<double> rr(K), MM((size_t)K * K), ls(K, 0.0), lp((size_t)K * K, 0.0);
            double ln = 0.0;
#ifdef _OPENMP
#pragma omp for schedule(guided)
#endif
            for (int i = 0; i < N; ++i) {
                if (failed) continue;
                try {
                    sample_item(rng, i, other, alpha, rr.data(), MM.data());
                } catch (const CholeskyFailed &) {
                    failed = true;
                    continue;
                }
                double *dst = &items[(size_t)i * K];
                for (int a = 0; a < K; ++a) dst[a] = rr[a];   // :324
                for (int b = 0; b < K; ++b)
                    for (int a = 0; a < K; ++a) lp[a + b * K] += rr[a] * rr[b];  // :359-360
                double sq = 0.0;                              // r.squaredNorm() is formed per item, then added (:362)
                for (int a = 0; a < K; ++a) { ls[a] += rr[a]; sq += rr[a] * rr[a]; }  // :361-362
                ln += sq;
                if (keep_aggr && iter >= burnin) {            // :364-368
                    for (int a = 0; a < K; ++a) aggrMu[(size_t)i * K + a] += rr[a];
                    double *al = &aggrLambda[(size_t)i * K * K];
                    for (int b = 0; b < K; ++b)
                        for (int a = 0; a < K; ++a) al[a + b * K] += rr[a] * rr[b];
                }
            }
#ifdef _OPENMP
#pragma omp critical
#endif
            {
                for (int a = 0; a < K; ++a) s[a] += ls[a];
                for (size_t e = 0; e < (size_t)K * K; ++e) prod[e] += lp[e];
                nrm += ln;
            }
        }
        if (failed) throw CholeskyFailed();
        norm = nrm;  // :381
        for (int b = 0; b < K; ++b)  // :383-384 ; the LOCAL sum is used, the member stays 0 (Q1)
            for (int a = 0; a < K; ++a) cov[a + b * K] = (prod[a + b * K] - (s[a] * s[b] / N)) / (N - 1);
        last_sum = s; last_prod = prod;
    }

    // sample.cpp:48-96 (NO_COMM: lo=0, hi=num() either way)
    void predict(const Side &other, int burnin)
    {
        const int n = (iter < burnin) ? 0 : (iter - burnin);
        double se = 0.0, se_avg = 0.0;
        int64_t nump = 0;
        for (int k = 0; k < T.ncols; ++k)
            for (int64_t p = T.colptr[k]; p < T.colptr[k + 1]; ++p) {
                const double *m = &items[(size_t)k * K];
                const double *u = &other.items[(size_t)T.rowidx[p] * K];
                double d = 0.0;
                for (int a = 0; a < K; ++a) d += m[a] * u[a];
                const double pred = d + mean_rating;
                se += (T.val[p] - pred) * (T.val[p] - pred);
                double &avg = Pavg[p];
                const double delta = pred - avg;
                avg = (n == 0) ? pred : (avg + delta / n);
                double &m2 = Pm2[p];
                m2 = (n == 0) ? 0 : m2 + delta * (pred - avg);
                se_avg += (T.val[p] - avg) * (T.val[p] - avg);
                nump++;
            }
        num_predict = nump;
        rmse = std::sqrt(se / (double)nump);
        rmse_avg = std::sqrt(se_avg / (double)nump);
    }
};

// ----------------------------------------------------------------------------------------------
// The two factors + the main-loop order of bpmf.cpp
// ----------------------------------------------------------------------------------------------
struct Model {
    int K;
    double alpha = 2.0;  // sample.cpp:29
    int burnin = 5;      // bpmf.cpp:79
    int nthreads = 0;
    Side movies, users;

    // file rows = users, file columns = movies (sample.cpp:112-137)
    Model(int K_, int nrows, int ncols, int64_t nnz, const int32_t *r, const int32_t *c, const double *v,
          int64_t nnz_t, const int32_t *tr, const int32_t *tc, const double *tv, int trows, int tcols,
          bool keep_aggr, bool no_cov)
        : K(K_)
    {
        movies.name = "movs"; users.name = "users";
        movies.M = Csc::from_coo(nrows, ncols, nnz, r, c, v);
        movies.T = Csc::from_coo(trows, tcols, nnz_t, tr, tc, tv);
        const int R = std::max(nrows, trows), C = std::max(ncols, tcols);
        movies.M.resize_dims(R, C);
        movies.T.resize_dims(R, C);
        users.M = movies.M.transpose();
        users.T = movies.T.transpose();
        movies.keep_aggr = users.keep_aggr = keep_aggr;
        movies.no_covariance = users.no_covariance = no_cov;
        movies.init(K); users.init(K);
    }
    void iterate()  // bpmf.cpp:184-190
    {
        movies.sweep(users, alpha, burnin, nthreads);
        users.sweep(movies, alpha, burnin, nthreads);
        movies.predict(users, burnin);
        users.predict(movies, burnin);
    }
    void finish() { movies.predict(users, burnin); }  // bpmf.cpp:225|242 (quirk Q4)
};

}  // namespace bpmf_oracle
