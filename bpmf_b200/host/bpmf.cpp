// bpmf.cpp — the `bpmf` executable: same command line, same start-up report, same per-iteration line and the same
// output files as the reference's driver (c++/bpmf.cpp:41-260), with the sampling loop running on B200s through
// CUDA_Sys (cuda_sys.h -> libbpmf_b200.so). K is a run-time option here (-d K, default 32; the reference compiles
// one binary per K, c++/bpmf.h:53) and -g N chooses the number of GPUs of this box to use.
#include <unistd.h>

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "io.h"
#include "sys.h"

#include "cuda_sys.h"   // the back end: defines SYS

using bpmf_host::write_matrix;

static void usage()
{
    std::cout << "Usage: bpmf -n <MTX> -p <MTX> [-o DIR/] [-i N] [-b N] [-f N] [-a F] [-d K] [-g N] [-krv] [-t N]\n"
              << "\n"
              << "Paramaters: \n"
              << "  -n MTX: Training input data\n"
              << "  -p MTX: Test input data\n"
              << "  [-o DIR]: Output directory for model and predictions\n"
              << "  [-i N]: Number of total iterations\n"
              << "  [-b N]: Number of burnin iterations\n"
              << "  [-f N]: Frequency to send model other nodes (accepted, unused — as in the reference)\n"
              << "  [-a F]: Noise precision (alpha)\n"
              << "  [-d K]: Number of latent dimensions (default 32)\n"
              << "  [-g N]: Number of GPUs of this box to use (default 1)\n"
              << "\n"
              << "  [-k]: Do not optimize item to node assignment (accepted, no effect with one process)\n"
              << "  [-r]: Redirect stdout to file\n"
              << "  [-v]: Output all samples\n"
              << "  [-t N]: Number of OpenMP threads (accepted, unused: the sweep runs on the GPU)\n"
              << "  [-x]: use the reference-order (exact) kernel instead of the fastest one\n"
              << "\n"
              << "Matrix Formats:\n"
              << "  *.mtx: Sparse or dense Matrix Market format\n"
              << "  *.sdm: Sparse binary double format\n"
              << "  *.ddm: Dense binary double format\n"
              << std::endl;
}

static int run(int argc, char *argv[])
{
    int ch;
    std::string fname, probename;
    std::string mname, lname;
    int nthrds = -1;
    bool redirect = false;
    Sys::nsims = 20;
    Sys::burnin = 5;
    Sys::update_freq = 1;

    // the reference's option string (c++/bpmf.cpp:83) plus -x; -g is in the reference's string but unused there
    while ((ch = getopt(argc, argv, "krvxn:t:p:i:b:f:g:w:u:o:s:m:l:a:d:")) != -1) {
        switch (ch) {
            case 'i': Sys::nsims = atoi(optarg); break;
            case 'b': Sys::burnin = atoi(optarg); break;
            case 'f': Sys::update_freq = atoi(optarg); break;
            case 't': nthrds = atoi(optarg); break;
            case 'a': Sys::alpha = atof(optarg); break;
            case 'd': num_latent = atoi(optarg); break;
            case 'g': CUDA_Sys::ngpus = atoi(optarg); break;
            case 'n': fname = optarg; break;
            case 'p': probename = optarg; break;
            case 'o': Sys::odirname = optarg; break;
            case 'm': mname = optarg; break;
            case 'l': lname = optarg; break;
            case 'r': redirect = true; break;
            case 'k': Sys::permute = false; break;
            case 'v': Sys::verbose = true; break;
            case 'x': CUDA_Sys::kernel_variant = BPMF_GPU_KERNEL_EXACT; break;
            case 'w': case 'u': case 's': break;
            case '?':
            case 'h':
            default: usage(); Sys::Abort(1);
        }
    }
    (void)nthrds;

    if (Sys::nprocs > 1 || redirect) {
        std::stringstream ofname;
        ofname << "bpmf_" << Sys::procid << ".out";
        Sys::os = new std::ofstream(ofname.str());
    } else {
        Sys::os = &std::cout;
    }
    Sys::dbgs = new std::ofstream("/dev/null");

    if (fname.empty() || probename.empty() || num_latent < 1 || num_latent > 128 || CUDA_Sys::ngpus < 1) {
        usage();
        Sys::Abort(1);
    }

    SYS movies("movs", fname, probename);
    SYS users("users", movies.M, movies.Pavg);

    movies.add_prop_posterior(mname);
    users.add_prop_posterior(lname);

    movies.alloc_and_init();
    users.alloc_and_init();

    movies.assign(users);
    users.assign(movies);
    users.build_conn(movies);
    movies.build_conn(users);

    long double average_items_sec = .0;
    long double average_ratings_sec = .0;

    char name[1024];
    gethostname(name, 1024);
    Sys::cout() << "hostname: " << name << std::endl;
    Sys::cout() << "pid: " << getpid() << std::endl;
    if (getenv("PBS_JOBID")) Sys::cout() << "jobid: " << getenv("PBS_JOBID") << std::endl;

    if (Sys::procid == 0) {
        Sys::cout() << "num_latent: " << num_latent << std::endl;
        Sys::cout() << "nprocs: " << Sys::nprocs << std::endl;
        Sys::cout() << "nthrds: " << 1 << std::endl;
        Sys::cout() << "ngpus: " << CUDA_Sys::ngpus << std::endl;
        Sys::cout() << "nsims: " << Sys::nsims << std::endl;
        Sys::cout() << "burnin: " << Sys::burnin << std::endl;
        Sys::cout() << "alpha: " << Sys::alpha << std::endl;
        Sys::cout() << "update_freq: " << Sys::update_freq << std::endl;
    }

    Sys::sync();

    auto begin = tick();

    for (int i = 0; i < Sys::nsims; ++i) {
        auto start = tick();

        movies.sample(users);
        users.sample(movies);

        movies.predict(users);
        users.predict(movies);

        auto stop = tick();
        double items_per_sec = (users.num() + movies.num()) / (stop - start);
        double ratings_per_sec = (users.nnz()) / (stop - start);
        movies.print(items_per_sec, ratings_per_sec, sqrt(users.norm), sqrt(movies.norm));
        average_items_sec += items_per_sec;
        average_ratings_sec += ratings_per_sec;

        if (Sys::verbose) {
            users.bcast();
            movies.bcast();
            if (Sys::procid == 0) {
                write_matrix(Sys::odirname + "/U-" + std::to_string(i) + ".ddm", users.items(), num_latent, users.num());
                write_matrix(Sys::odirname + "/V-" + std::to_string(i) + ".ddm", movies.items(), num_latent, movies.num());
            }
        }
    }

    Sys::sync();

    auto end = tick();
    auto elapsed = end - begin;

    users.bcast();
    movies.bcast();

    if (Sys::odirname.size()) {
        users.unpermuteCols(movies);
        movies.unpermuteCols(users);
        movies.predict(users, true);

        if (Sys::procid == 0) {
            // sparse
            write_matrix(Sys::odirname + "/Pavg.sdm", movies.Pavg);
            write_matrix(Sys::odirname + "/Pm2.sdm", movies.Pm2);

            // dense
            users.finalize_mu_lambda();
            write_matrix(Sys::odirname + "/U-mu.ddm", users.aggrMu);
            write_matrix(Sys::odirname + "/U-Lambda.ddm", users.aggrLambda);

            movies.finalize_mu_lambda();
            write_matrix(Sys::odirname + "/V-mu.ddm", movies.aggrMu);
            write_matrix(Sys::odirname + "/V-Lambda.ddm", movies.aggrLambda);
        }
    } else {
        movies.predict(users, true);
    }

    if (Sys::procid == 0) {
        Sys::cout() << "Total time: " << elapsed << std::endl << std::flush;
        Sys::cout() << "Final Avg RMSE: " << movies.rmse_avg << std::endl << std::flush;
        Sys::cout() << "  computed on " << movies.num_predict << " items ("
                    << int(100. * movies.num_predict / movies.T.nonZeros()) << "% of total items in test set)" << std::endl
                    << std::flush;
        Sys::cout() << "Average items/sec: " << average_items_sec / movies.iter << std::endl << std::flush;
        Sys::cout() << "Average ratings/sec: " << average_ratings_sec / movies.iter << std::endl << std::flush;
    }
    return 0;
}

int main(int argc, char *argv[])
{
    Sys::Init();
    int rc = run(argc, argv);   // exceptions are not caught: std::terminate, as in the reference (c++/error.h)
    Sys::Finalize();
    return rc;
}
