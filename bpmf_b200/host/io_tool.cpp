// io_tool — converts between the matrix file formats of io.h, the extension decides (c++/io.cpp:31-77):
//     io_tool sparse <in> <out>      read_matrix(SparseMatrixD) -> write_matrix
//     io_tool dense  <in> <out>      read_matrix(DenseMatrixD)  -> write_matrix
//     io_tool info   <in>            prints "<rows> <cols> <nnz> <sum>" of a sparse file
// Used by tests/test_host_io.py to check the loaders against numpy/scipy, and handy for turning a text .mtx into .sdm.
#include <cstdio>
#include <cstring>
#include <exception>

#include "io.h"

int main(int argc, char **argv)
{
    using namespace bpmf_host;
    try {
        if (argc == 4 && !strcmp(argv[1], "sparse")) {
            SparseMatrixD X;
            read_matrix(argv[2], X);
            write_matrix(argv[3], X);
            return 0;
        }
        if (argc == 4 && !strcmp(argv[1], "dense")) {
            DenseMatrixD X;
            read_matrix(argv[2], X);
            write_matrix(argv[3], X);
            return 0;
        }
        if (argc == 3 && !strcmp(argv[1], "info")) {
            SparseMatrixD X;
            read_matrix(argv[2], X);
            printf("%lld %lld %lld %.17g\n", (long long)X.rows(), (long long)X.cols(), (long long)X.nonZeros(), X.sum());
            return 0;
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "io_tool: %s\n", e.what());
        return 2;
    }
    fprintf(stderr, "usage: io_tool sparse|dense <in> <out> | io_tool info <in>\n");
    return 1;
}
