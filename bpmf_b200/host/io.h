// io.h — matrix file loaders / writers with the reference's formats and entry-point names
// (c++/io.h:36-60): the format is chosen by the file extension, ".gz" on top of any of them.
//   .mtx .mm  Matrix Market text   (coordinate real|pattern general -> sparse, array real general -> dense)
//   .sdm      sparse binary fp64   u64 nrow, ncol, nnz; u32 rows[nnz] (1-based); u32 cols[nnz]; f64 vals[nnz]
//   .sbm      sparse binary pattern (as .sdm without the values; every entry 1.0)
//   .ddm      dense binary fp64    u64 nrow, ncol; f64 data[nrow*ncol] column-major
//   .csv      dense text           nrow \n ncol \n comma-separated rows
// Errors are std::runtime_error, as THROWERROR in the reference (c++/error.h:18-30).
#pragma once
#include <string>

#include "matrix.h"

namespace bpmf_host {

struct MatrixType {
    enum Kind { none, sdm, sbm, mtx, csv, ddm } type;
    bool compressed;
};
MatrixType ExtensionToMatrixType(const std::string &fname);

void read_matrix(const std::string &filename, SparseMatrixD &X);
void read_matrix(const std::string &filename, DenseMatrixD &X);
// the entries of a sparse matrix file in file order, without building the compressed form (same formats, same refusals)
void read_matrix(const std::string &filename, TripletList &X);
void write_matrix(const std::string &filename, const SparseMatrixD &X);
void write_matrix(const std::string &filename, const DenseMatrixD &X);
// a K x N latent matrix that lives in a raw buffer (Sys::items_ptr)
void write_matrix(const std::string &filename, const double *colmajor, int64_t nrows, int64_t ncols);

}  // namespace bpmf_host
