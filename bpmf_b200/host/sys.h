// sys.h — `struct Sys`, one per factor ("movies" or "users"), with the reference's member names and call
// contract (c++/bpmf.h:112-239) but no Eigen types: matrices are bpmf_host::SparseMatrixD / DenseMatrixD and
// the latent matrix is the raw `items_ptr` (K x num(), item i at items_ptr + i*K, c++/bpmf.h:193-194).
//
// A back end is a header that `#define SYS <Derived>`, derives from Sys with the two constructors, overrides
// alloc_and_init() / send_item() / sample() and defines the statics Init/Finalize/sync/Abort — the contract
// c++/nocomm.h:6-37 fulfils for NO_COMM. cuda_sys.h is that header for the B200.
//
// There is deliberately NO host implementation of the numerical path here: Sys::sample and Sys::predict of
// the base class throw. Only a device back end can run them.
#pragma once
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "matrix.h"

using bpmf_host::DenseMatrixD;
using bpmf_host::SparseMatrixD;

#define THROWERROR(msg)                                                                                          \
    throw std::runtime_error(std::string("line: ") + std::to_string(__LINE__) + " file: " + __FILE__ + " function: " + \
                             __func__ + "\n" + (msg))

extern int num_latent;   // K: a run-time value here (-d K), a compile-time constant in the reference (c++/bpmf.h:53)

const int breakpoint1 = 24;      // c++/bpmf.h:255-256
const int breakpoint2 = 10500;

double tick();                   // wall-clock seconds (c++/counters.cpp:160-163)

// host mirror of c++/bpmf.h:78-104; the draw itself (CondNormalWishart) runs on the device
struct HyperParams {
    std::vector<double> mu, LambdaF, LambdaU;   // K, K*K col-major, K*K upper
    void resize(int K)
    {
        mu.assign((size_t)K, 0.0);
        LambdaF.assign((size_t)K * K, 0.0);
        LambdaU.assign((size_t)K * K, 0.0);
    }
};

struct Sys {
    //-- static info (c++/bpmf.h:114-138)
    static bool permute;
    static bool verbose;
    static int nprocs, procid;
    static int burnin, nsims, update_freq;
    static double alpha;
    static std::string odirname;

    static void Init();
    static void Finalize();
    static void Abort(int);
    static void sync();

    static std::ostream *os, *dbgs;
    static std::ostream &cout()
    {
        if (!os) return std::cout;
        os->flush();
        return *os;
    }
    static std::ostream &dbg()
    {
        if (!dbgs) return std::cerr;
        dbgs->flush();
        return *dbgs;
    }

    //-- c'tor
    std::string name;
    int iter;
    Sys(std::string name, std::string fname, std::string pname);
    Sys(std::string name, const SparseMatrixD &M, const SparseMatrixD &Pavg);
    virtual ~Sys();

  protected:
    // for a back end that fills M / T itself (cuda_sys.h builds both factors' matrices on the device): members
    // initialised, nothing loaded
    explicit Sys(std::string name);

  public:
    void init();
    virtual void alloc_and_init() = 0;

    //-- sparse matrix
    SparseMatrixD M;   // known ratings, column = item of this factor
    double mean_rating;
    int num() const { return (int)M.cols(); }
    long nnz() const { return (long)M.nonZeros(); }
    int nnz(int i) const { return (int)M.col_nnz(i); }

    // assignment of items to nodes: contiguous, node i owns [from(i), to(i))
    void assign(Sys &);
    void build_conn(Sys &) {}            // one process: nothing to connect (c++/assign.cpp:281)
    void unpermuteCols(Sys &) {}         // no permutation is ever applied here
    bool assigned;
    std::vector<int> dom;
    int num(int i) const { return to(i) - from(i); }
    int from(int i = procid) const { return dom.at((size_t)i); }
    int to(int i = procid) const { return dom.at((size_t)i + 1); }

    //-- factors of the MF
    double *items_ptr;
    const double *items() const { return items_ptr; }
    const double *items_col(long i) const { return items_ptr + i * num_latent; }

    //-- propagated posterior (-m / -l): per-item priors, uploaded by the back end
    DenseMatrixD propMu, propLambda;
    void add_prop_posterior(std::string);
    bool has_prop_posterior() const { return propMu.nonZeros() > 0; }

    //-- aggregated posterior (-o)
    DenseMatrixD aggrMu, aggrLambda;
    virtual void finalize_mu_lambda();

    // overridden by the back end
    virtual void send_item(int i) = 0;
    void bcast() {}                      // one process owns everything (c++/bpmf.cpp:263-278 asserts nprocs == 1)
    virtual void sample(Sys &in);

    std::vector<double> sum;   //-- never updated by the reference either
    std::vector<double> cov;   //-- K x K covariance of the last sweep
    double norm;

    //-- hyper params
    HyperParams hp;

    // output predictions
    SparseMatrixD T, Torig;    // test matrix (input)
    SparseMatrixD Pavg, Pm2;   // predictions for items in T (output)
    double rmse, rmse_avg;
    long num_predict;
    virtual void predict(Sys &other, bool all = false);
    void print(double, double, double, double);
};
