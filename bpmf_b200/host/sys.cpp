// sys.cpp — the back-end independent part of `struct Sys` (see sys.h): statics, the two constructors, init() with
// the reference's start-up report, the per-iteration print line, the -o post-processing. Mirrors the behaviour of
// c++/sample.cpp:20-226 and c++/bpmf.cpp:281-295 with Eigen-free containers.
//
// The numerical path is NOT here: Sys::sample / Sys::predict throw unless a device back end (cuda_sys.h) overrides
// them. That is deliberate — there is no CPU fallback in this product.
#include "sys.h"

#include <chrono>
#include <cmath>
#include <cstdio>

#include "io.h"

int num_latent = 32;

std::ostream *Sys::os = nullptr;
std::ostream *Sys::dbgs = nullptr;

int Sys::procid = -1;
int Sys::nprocs = -1;

int Sys::nsims = 20;          // c++/bpmf.cpp:78-80
int Sys::burnin = 5;
int Sys::update_freq = 1;
double Sys::alpha = 2.0;      // c++/sample.cpp:29

std::string Sys::odirname = "";

bool Sys::permute = true;
bool Sys::verbose = false;

double tick()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

// c++/sample.cpp:112-127: train + test from files, both grown to the common shape, Pm2 = Pavg = Torig = T
Sys::Sys(std::string name_, std::string fname, std::string probename)
    : name(std::move(name_)), iter(-1), mean_rating(0.0), assigned(false), dom((size_t)(nprocs > 0 ? nprocs : 1) + 1, 0),
      items_ptr(nullptr), norm(0.0), rmse(0.0), rmse_avg(0.0), num_predict(0)
{
    bpmf_host::read_matrix(fname, M);
    bpmf_host::read_matrix(probename, T);
    const int64_t rows = std::max(M.rows(), T.rows());
    const int64_t cols = std::max(M.cols(), T.cols());
    M.conservativeResize(rows, cols);
    T.conservativeResize(rows, cols);
    Pm2 = Pavg = Torig = T;
}

Sys::Sys(std::string name_)
    : name(std::move(name_)), iter(-1), mean_rating(0.0), assigned(false), dom((size_t)(nprocs > 0 ? nprocs : 1) + 1, 0),
      items_ptr(nullptr), norm(0.0), rmse(0.0), rmse_avg(0.0), num_predict(0)
{
}

// c++/sample.cpp:132-137: the other factor is the transpose of an existing one
Sys::Sys(std::string name_, const SparseMatrixD &Mt, const SparseMatrixD &Pt)
    : name(std::move(name_)), iter(-1), mean_rating(0.0), assigned(false), dom((size_t)(nprocs > 0 ? nprocs : 1) + 1, 0),
      items_ptr(nullptr), norm(0.0), rmse(0.0), rmse_avg(0.0), num_predict(0)
{
    M = Mt.transpose();
    Pm2 = Pavg = T = Torig = Pt.transpose();
    if (M.rows() != Pavg.rows() || M.cols() != Pavg.cols()) THROWERROR("train and test matrices differ in shape");
}

Sys::~Sys() {}

// c++/sample.cpp:157-174: the propagated posterior of an earlier run (its U-mu / U-Lambda outputs) as per-item priors.
// The back end uploads propLambda in alloc_and_init (bpmf_gpu_set_prop_posterior).
void Sys::add_prop_posterior(std::string fnames)
{
    if (fnames.empty()) return;
    const std::size_t pos = fnames.find_first_of(",");
    const std::string mu_name = fnames.substr(0, pos);
    const std::string lambda_name = fnames.substr(pos + 1);
    bpmf_host::read_matrix(mu_name, propMu);
    bpmf_host::read_matrix(lambda_name, propLambda);
    if (propMu.cols() != num() || propLambda.cols() != num()) THROWERROR("propagated posterior: wrong number of columns");
    if (propMu.rows() != num_latent || propLambda.rows() != (int64_t)num_latent * num_latent)
        THROWERROR("propagated posterior: wrong number of rows");
}

// c++/sample.cpp:179-226 (items().setZero() happens in the back end, which owns items_ptr)
void Sys::init()
{
    if (!(M.rows() > 0 && M.cols() > 0)) THROWERROR("empty train matrix");
    mean_rating = M.sum() / (double)M.nonZeros();
    const size_t K = (size_t)num_latent;
    if (items_ptr) std::fill(items_ptr, items_ptr + K * (size_t)num(), 0.0);
    sum.assign(K, 0.0);
    cov.assign(K * K, 0.0);
    norm = 0.0;
    hp.resize(num_latent);

    if (Sys::odirname.size()) {
        aggrMu.resize(num_latent, num());
        aggrLambda.resize((int64_t)num_latent * num_latent, num());
    }

    long count_larger_bp1 = 0, count_larger_bp2 = 0, count_sum = 0;
    for (int k = 0; k < num(); k++) {
        const int count = nnz(k);
        count_sum += count;
        if (count > breakpoint1) count_larger_bp1++;
        if (count > breakpoint2) count_larger_bp2++;
    }

    Sys::cout() << "mean rating: " << mean_rating << std::endl;
    Sys::cout() << "total number of ratings in train: " << M.nonZeros() << std::endl;
    Sys::cout() << "total number of ratings in test: " << T.nonZeros() << std::endl;
    Sys::cout() << "average ratings per row: " << (double)count_sum / (double)M.cols() << std::endl;
    Sys::cout() << "rows > break_point1: " << 100. * (double)count_larger_bp1 / (double)M.cols() << std::endl;
    Sys::cout() << "rows > break_point2: " << 100. * (double)count_larger_bp2 / (double)M.cols() << std::endl;
    Sys::cout() << "num " << name << ": " << num() << std::endl;
    if (has_prop_posterior()) Sys::cout() << "with propagated posterior" << std::endl;
}

// one process: every item is local (c++/assign.cpp:54-58 for nprocs == 1)
void Sys::assign(Sys &)
{
    dom.assign((size_t)nprocs + 1, 0);
    for (int p = 1; p <= nprocs; ++p) dom[(size_t)p] = num();
    assigned = true;
}

// c++/sample.cpp:101-107
void Sys::print(double items_per_sec, double ratings_per_sec, double norm_u, double norm_m)
{
    char buf[1024];
    const char *phase = (iter < Sys::burnin) ? "Burnin" : "Sampling";
    snprintf(buf, sizeof buf,
             "%d: %s iteration %d:\t RMSE: %3.4f\tavg RMSE: %3.4f\tFU(%6.2f)\tFM(%6.2f)\titems/sec: %6.2f\tratings/sec: %6.2fM\n",
             Sys::procid, phase, iter, rmse, rmse_avg, norm_u, norm_m, items_per_sec, ratings_per_sec / 1e6);
    Sys::cout() << buf;
}

void Sys::sample(Sys &) { THROWERROR("Sys::sample has no host implementation: use a device back end (cuda_sys.h)"); }
void Sys::predict(Sys &, bool) { THROWERROR("Sys::predict has no host implementation: use a device back end (cuda_sys.h)"); }

namespace {

// general inverse by LU with partial pivoting (what Eigen's .inverse() does for K > 4), column-major K x K
void inverse_colmajor(std::vector<double> &A, int K, std::vector<double> &inv)
{
    std::vector<int> piv((size_t)K);
    auto a = [&](int r, int c) -> double & { return A[(size_t)r + (size_t)c * K]; };
    for (int k = 0; k < K; ++k) {
        int p = k;
        double best = std::fabs(a(k, k));
        for (int r = k + 1; r < K; ++r)
            if (std::fabs(a(r, k)) > best) { best = std::fabs(a(r, k)); p = r; }
        piv[(size_t)k] = p;
        if (p != k)
            for (int c = 0; c < K; ++c) std::swap(a(k, c), a(p, c));
        const double d = a(k, k);
        for (int r = k + 1; r < K; ++r) a(r, k) /= d;
        for (int c = k + 1; c < K; ++c) {
            const double u = a(k, c);
            for (int r = k + 1; r < K; ++r) a(r, c) -= a(r, k) * u;
        }
    }
    inv.assign((size_t)K * K, 0.0);
    std::vector<double> b((size_t)K);
    for (int c = 0; c < K; ++c) {
        for (int r = 0; r < K; ++r) b[(size_t)r] = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < K; ++k) std::swap(b[(size_t)k], b[(size_t)piv[(size_t)k]]);
        for (int k = 0; k < K; ++k)
            for (int r = k + 1; r < K; ++r) b[(size_t)r] -= a(r, k) * b[(size_t)k];
        for (int k = K - 1; k >= 0; --k) {
            b[(size_t)k] /= a(k, k);
            for (int r = 0; r < k; ++r) b[(size_t)r] -= a(r, k) * b[(size_t)k];
        }
        for (int r = 0; r < K; ++r) inv[(size_t)r + (size_t)c * K] = b[(size_t)r];
    }
}

}  // namespace

// c++/bpmf.cpp:281-295: per item, covariance of the post-burn-in samples -> precision; mean. Output post-processing,
// run once after the loop on the aggregates the device accumulated.
void Sys::finalize_mu_lambda()
{
    if (!aggrLambda.nonZeros() || !aggrMu.nonZeros()) THROWERROR("no aggregated posterior: run with -o");
    const int K = num_latent;
    const int nsamples = Sys::nsims - Sys::burnin;
    std::vector<double> cv((size_t)K * K), prec;
    for (int i = 0; i < num(); i++) {
        double *s = aggrMu.col(i);
        double *prod = aggrLambda.col(i);
        for (int c = 0; c < K; ++c)
            for (int r = 0; r < K; ++r)
                cv[(size_t)r + (size_t)c * K] = (prod[(size_t)r + (size_t)c * K] - (s[r] * s[c] / nsamples)) / (nsamples - 1);
        inverse_colmajor(cv, K, prec);
        std::copy(prec.begin(), prec.end(), prod);
        for (int r = 0; r < K; ++r) s[r] = s[r] / nsamples;
    }
}
