// ref_io_harness.cpp — TEST INFRASTRUCTURE. The reference's file-format code (c++/io.cpp + c++/gzstream.cpp, compiled
// unmodified from /root/reference against the stand-in Eigen headers of oracle/shim/, oracle/Makefile target `ref`) behind
// the same three operations as bpmf_b200/host/io_tool, so that tests/test_host_io_vs_reference.py can hold the host
// loaders / writers of bpmf_b200/host/io.cpp to it file by file.
#include <cstdio>
#include <exception>
#include <string>

#include "io.h"

namespace {
std::string last_error;
}

extern "C" {

const char *bpmf_ref_io_error() { return last_error.c_str(); }

// read_matrix(in) -> write_matrix(out); dense != 0: Eigen::MatrixXd, else Eigen::SparseMatrix<double>. 0 = ok.
int bpmf_ref_io_convert(const char *in, const char *out, int dense)
{
    try {
        if (dense) {
            Eigen::MatrixXd X;
            read_matrix(in, X);
            write_matrix(out, X);
        } else {
            Eigen::SparseMatrix<double> X;
            read_matrix(in, X);
            write_matrix(out, X);
        }
    } catch (const std::exception &e) {
        last_error = e.what();
        return 1;
    }
    return 0;
}

// rows, cols, nnz and the sum of the stored values of a sparse file. 0 = ok.
int bpmf_ref_io_info(const char *in, long long *rows, long long *cols, long long *nnz, double *sum)
{
    try {
        Eigen::SparseMatrix<double> X;
        read_matrix(in, X);
        *rows = X.rows(); *cols = X.cols(); *nnz = X.nonZeros(); *sum = X.sum();
    } catch (const std::exception &e) {
        last_error = e.what();
        return 1;
    }
    return 0;
}

}  // extern "C"
