// bpmf.cpp — the `bpmf` executable. A drop-in for the reference's driver (c++/bpmf.cpp:41-260): it takes the same command
// line, prints the same start-up report, per-iteration line and closing summary, and writes the same output files, while
// the sampling runs on B200s through CUDA_Sys (cuda_sys.h -> libbpmf_b200.so). Differences: K is a run-time option
// (-d K, default 32; the reference compiles one binary per K, c++/bpmf.h:53), -g N uses N GPUs of this box, and -x
// selects the reference-order kernel.
#include <unistd.h>

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "io.h"
#include "sys.h"

#include "cuda_sys.h"   // the back end: defines SYS

namespace {

struct Options {
    std::string train, test;          // -n, -p
    std::string prior_movies, prior_users;   // -m, -l: "<mu file>,<Lambda file>" of an earlier run
    bool to_file = false;             // -r
    bool bad = false;
};

// One row per option of the reference's getopt string (c++/bpmf.cpp:83, "krvn:t:p:i:b:f:g:w:u:o:s:m:l:a:" there) plus
// -d and -x; `apply` receives the argument text (nullptr for flags).
struct OptionRow {
    char letter;
    bool has_arg;
    const char *help;                 // nullptr: accepted and ignored, not listed
    void (*apply)(Options &, const char *);
};

const OptionRow OPTION_TABLE[] = {
    {'n', true, "  -n MTX: Training input data", [](Options &o, const char *a) { o.train = a; }},
    {'p', true, "  -p MTX: Test input data", [](Options &o, const char *a) { o.test = a; }},
    {'o', true, "  [-o DIR]: Output directory for model and predictions", [](Options &, const char *a) { Sys::odirname = a; }},
    {'i', true, "  [-i N]: Number of total iterations", [](Options &, const char *a) { Sys::nsims = atoi(a); }},
    {'b', true, "  [-b N]: Number of burnin iterations", [](Options &, const char *a) { Sys::burnin = atoi(a); }},
    {'f', true, "  [-f N]: Frequency to send model other nodes (accepted, unused — as in the reference)",
     [](Options &, const char *a) { Sys::update_freq = atoi(a); }},
    {'a', true, "  [-a F]: Noise precision (alpha)", [](Options &, const char *a) { Sys::alpha = atof(a); }},
    {'d', true, "  [-d K]: Number of latent dimensions (default 32)", [](Options &, const char *a) { num_latent = atoi(a); }},
    {'g', true, "  [-g N]: Number of GPUs of this box to use (default 1)", [](Options &, const char *a) { CUDA_Sys::ngpus = atoi(a); }},
    {'m', true, "  [-m MU,LAMBDA]: propagated posterior of the movies (U-mu / U-Lambda files of an earlier run)",
     [](Options &o, const char *a) { o.prior_movies = a; }},
    {'l', true, "  [-l MU,LAMBDA]: propagated posterior of the users", [](Options &o, const char *a) { o.prior_users = a; }},
    {'k', false, "\n  [-k]: Do not optimize item to node assignment (accepted, no effect with one process)",
     [](Options &, const char *) { Sys::permute = false; }},
    {'r', false, "  [-r]: Redirect stdout to file", [](Options &o, const char *) { o.to_file = true; }},
    {'v', false, "  [-v]: Output all samples", [](Options &, const char *) { Sys::verbose = true; }},
    {'t', true, "  [-t N]: Number of OpenMP threads (accepted, unused: the sweep runs on the GPU)", [](Options &, const char *) {}},
    {'x', false, "  [-x]: use the reference-order (exact) kernel instead of the fastest one",
     [](Options &, const char *) { CUDA_Sys::kernel_variant = BPMF_GPU_KERNEL_EXACT; }},
    {'H', false, "  [-H]: build the compressed matrices on the host (default: on the GPU, from the files' entry lists)",
     [](Options &, const char *) { CUDA_Sys::device_build = false; }},
    {'w', true, nullptr, [](Options &, const char *) {}},
    {'u', true, nullptr, [](Options &, const char *) {}},
    {'s', true, nullptr, [](Options &, const char *) {}},
};

void usage()
{
    std::cout << "Usage: bpmf -n <MTX> -p <MTX> [-o DIR/] [-i N] [-b N] [-f N] [-a F] [-d K] [-g N] [-krvxH] [-t N]\n\nParamaters: \n";
    for (const OptionRow &row : OPTION_TABLE)
        if (row.help) std::cout << row.help << "\n";
    std::cout << "\nMatrix Formats:\n"
              << "  *.mtx: Sparse or dense Matrix Market format\n"
              << "  *.sdm: Sparse binary double format\n"
              << "  *.ddm: Dense binary double format\n"
              << std::endl;
}

Options parse(int argc, char *argv[])
{
    Options opt;
    std::string spec;
    for (const OptionRow &row : OPTION_TABLE) {
        spec += row.letter;
        if (row.has_arg) spec += ':';
    }
    for (int ch; (ch = getopt(argc, argv, spec.c_str())) != -1;) {
        const OptionRow *hit = nullptr;
        for (const OptionRow &row : OPTION_TABLE)
            if (row.letter == ch) hit = &row;
        if (!hit) { opt.bad = true; break; }          // '?', -h, anything unknown
        hit->apply(opt, hit->has_arg ? optarg : nullptr);
    }
    if (opt.train.empty() || opt.test.empty() || num_latent < 1 || num_latent > 128 || CUDA_Sys::ngpus < 1) opt.bad = true;
    return opt;
}

// "key: value" lines of the start-up report (c++/bpmf.cpp:158-173)
template <typename T>
void report(const char *key, const T &value) { Sys::cout() << key << ": " << value << std::endl; }

void report_setup()
{
    char host[1024];
    gethostname(host, sizeof host);
    report("hostname", host);
    report("pid", getpid());
    if (const char *job = getenv("PBS_JOBID")) report("jobid", job);
    if (Sys::procid != 0) return;
    report("num_latent", num_latent);
    report("nprocs", Sys::nprocs);
    report("nthrds", 1);
    report("ngpus", CUDA_Sys::ngpus);
    report("nsims", Sys::nsims);
    report("burnin", Sys::burnin);
    report("alpha", Sys::alpha);
    report("update_freq", Sys::update_freq);
}

struct Throughput {
    long double items = 0, ratings = 0;   // sums of the per-iteration rates (the reference averages the rates)
};

// movies.sample(users); users.sample(movies); both predicts; the per-iteration line; -v dumps (c++/bpmf.cpp:180-210)
void gibbs_iteration(int it, SYS &movies, SYS &users, Throughput &acc)
{
    const double t0 = tick();
    movies.sample(users);
    users.sample(movies);
    movies.predict(users);
    users.predict(movies);
    const double dt = tick() - t0;
    const double items_rate = (users.num() + movies.num()) / dt, ratings_rate = users.nnz() / dt;
    movies.print(items_rate, ratings_rate, sqrt(users.norm), sqrt(movies.norm));
    acc.items += items_rate;
    acc.ratings += ratings_rate;
    if (!Sys::verbose) return;
    users.bcast();
    movies.bcast();
    if (Sys::procid != 0) return;
    const std::string tag = "-" + std::to_string(it) + ".ddm";
    bpmf_host::write_matrix(Sys::odirname + "/U" + tag, users.items(), num_latent, users.num());
    bpmf_host::write_matrix(Sys::odirname + "/V" + tag, movies.items(), num_latent, movies.num());
}

// -o: predictions and the posterior means / precisions of both factors (c++/bpmf.cpp:217-240)
void write_model(SYS &movies, SYS &users)
{
    users.unpermuteCols(movies);
    movies.unpermuteCols(users);
    movies.predict(users, true);
    if (Sys::procid != 0) return;
    const std::string &dir = Sys::odirname;
    bpmf_host::write_matrix(dir + "/Pavg.sdm", movies.Pavg);
    bpmf_host::write_matrix(dir + "/Pm2.sdm", movies.Pm2);
    struct { SYS *sys; const char *prefix; } factors[] = {{&users, "/U"}, {&movies, "/V"}};
    for (auto &f : factors) {
        f.sys->finalize_mu_lambda();
        bpmf_host::write_matrix(dir + f.prefix + "-mu.ddm", f.sys->aggrMu);
        bpmf_host::write_matrix(dir + f.prefix + "-Lambda.ddm", f.sys->aggrLambda);
    }
}

void report_summary(double elapsed, const SYS &movies, const Throughput &acc)
{
    if (Sys::procid != 0) return;
    std::ostream &out = Sys::cout();
    out << "Total time: " << elapsed << std::endl;
    out << "Final Avg RMSE: " << movies.rmse_avg << std::endl;
    out << "  computed on " << movies.num_predict << " items (" << int(100. * movies.num_predict / movies.T.nonZeros())
        << "% of total items in test set)" << std::endl;
    out << "Average items/sec: " << acc.items / movies.iter << std::endl;
    out << "Average ratings/sec: " << acc.ratings / movies.iter << std::endl;
}

int run(int argc, char *argv[])
{
    Sys::nsims = 20;                  // the reference's defaults (c++/bpmf.cpp:78-80)
    Sys::burnin = 5;
    Sys::update_freq = 1;
    const Options opt = parse(argc, argv);

    if (Sys::nprocs > 1 || opt.to_file) Sys::os = new std::ofstream("bpmf_" + std::to_string(Sys::procid) + ".out");
    else Sys::os = &std::cout;
    Sys::dbgs = new std::ofstream("/dev/null");
    if (opt.bad) {
        usage();
        Sys::Abort(1);
    }

    SYS movies("movs", opt.train, opt.test);          // columns of the train matrix
    SYS users("users", movies.M, movies.Pavg);        // its rows: the transposes
    movies.add_prop_posterior(opt.prior_movies);
    users.add_prop_posterior(opt.prior_users);
    for (SYS *s : {&movies, &users}) s->alloc_and_init();
    movies.assign(users);
    users.assign(movies);
    users.build_conn(movies);
    movies.build_conn(users);

    report_setup();
    Sys::sync();
    Throughput acc;
    const double begin = tick();
    for (int it = 0; it < Sys::nsims; ++it) gibbs_iteration(it, movies, users, acc);
    Sys::sync();
    const double elapsed = tick() - begin;

    users.bcast();
    movies.bcast();
    if (Sys::odirname.empty()) movies.predict(users, true);   // the final averaged prediction either way (quirk Q4)
    else write_model(movies, users);
    report_summary(elapsed, movies, acc);
    return 0;
}

}  // namespace

int main(int argc, char *argv[])
{
    Sys::Init();
    const int rc = run(argc, argv);   // exceptions are not caught: std::terminate, as in the reference (c++/error.h)
    Sys::Finalize();
    return rc;
}
