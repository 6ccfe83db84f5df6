// matrix.h — the two containers the host code needs, Eigen-free.
//
// SparseMatrixD keeps the semantics the reference gets from Eigen::SparseMatrix<double> (c++/bpmf.h:55):
// compressed columns, int32 inner (row) indices ascending inside a column, duplicates summed and explicit
// zeros kept by from_triplets (Eigen setFromTriplets, c++/io.cpp:282,521). DenseMatrixD is column-major
// like Eigen::MatrixXd, so a K x N latent matrix stores item i at data + i*K (c++/bpmf.h:193-194).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace bpmf_host {

struct DenseMatrixD {
    int64_t nrows = 0, ncols = 0;
    std::vector<double> v;
    DenseMatrixD() {}
    DenseMatrixD(int64_t r, int64_t c) : nrows(r), ncols(c), v((size_t)(r * c), 0.0) {}
    void resize(int64_t r, int64_t c) { nrows = r; ncols = c; v.assign((size_t)(r * c), 0.0); }
    int64_t rows() const { return nrows; }
    int64_t cols() const { return ncols; }
    double &operator()(int64_t r, int64_t c) { return v[(size_t)(r + c * nrows)]; }
    double operator()(int64_t r, int64_t c) const { return v[(size_t)(r + c * nrows)]; }
    double *data() { return v.data(); }
    const double *data() const { return v.data(); }
    double *col(int64_t c) { return v.data() + (size_t)(c * nrows); }
    const double *col(int64_t c) const { return v.data() + (size_t)(c * nrows); }
    int64_t nonZeros() const { return (int64_t)v.size(); }   // Eigen dense nonZeros() == size()
};

struct Triplet {
    int32_t row, col;
    double val;
};

// A matrix file as its readers see it: the (row, col, value) entries in FILE order. What bpmf_gpu_load_coo takes, so that the
// compressed matrices of both factors can be built on the device instead of on the host.
struct TripletList {
    int64_t nrows = 0, ncols = 0;
    std::vector<Triplet> t;
    bool refuse_duplicates = false;   // .sdm: the reference fails with "Invalid number of values" when entries coincide (c++/io.cpp:284-287)
    void from_triplets(int64_t nr, int64_t nc, std::vector<Triplet> &src) { nrows = nr; ncols = nc; t.swap(src); }
    int64_t nonZeros() const { return (int64_t)t.size(); }
};

struct SparseMatrixD {
    int64_t nrows = 0, ncols = 0;
    std::vector<int64_t> colptr{0};
    std::vector<int32_t> rowidx;
    std::vector<double> val;

    int64_t rows() const { return nrows; }
    int64_t cols() const { return ncols; }
    int64_t nonZeros() const { return (int64_t)val.size(); }
    int64_t col_nnz(int64_t c) const { return colptr[(size_t)c + 1] - colptr[(size_t)c]; }
    double sum() const
    {
        double s = 0.0;
        for (double x : val) s += x;
        return s;
    }

    // counting sort by column, then by row inside a column (stable, so duplicates are summed in input order)
    void from_triplets(int64_t nr, int64_t nc, const std::vector<Triplet> &t)
    {
        nrows = nr; ncols = nc;
        std::vector<int64_t> cnt((size_t)nc + 1, 0);
        for (const Triplet &e : t) {
            if (e.row < 0 || e.row >= nr || e.col < 0 || e.col >= nc) throw std::runtime_error("matrix entry out of range");
            cnt[(size_t)e.col + 1]++;
        }
        for (int64_t j = 0; j < nc; ++j) cnt[(size_t)j + 1] += cnt[(size_t)j];
        std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
        std::vector<int32_t> ri(t.size());
        std::vector<double> va(t.size());
        for (const Triplet &e : t) {
            const int64_t p = pos[(size_t)e.col]++;
            ri[(size_t)p] = e.row; va[(size_t)p] = e.val;
        }
        colptr.assign((size_t)nc + 1, 0);
        rowidx.clear(); val.clear();
        rowidx.reserve(t.size()); val.reserve(t.size());
        std::vector<std::pair<int32_t, int64_t>> ord;
        for (int64_t j = 0; j < nc; ++j) {
            const int64_t b = cnt[(size_t)j], e = cnt[(size_t)j + 1];
            bool sorted = true;
            for (int64_t p = b + 1; p < e && sorted; ++p) sorted = ri[(size_t)p - 1] < ri[(size_t)p];
            if (sorted) {
                rowidx.insert(rowidx.end(), ri.begin() + b, ri.begin() + e);
                val.insert(val.end(), va.begin() + b, va.begin() + e);
            } else {
                ord.clear();
                for (int64_t p = b; p < e; ++p) ord.emplace_back(ri[(size_t)p], p);
                std::stable_sort(ord.begin(), ord.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
                const size_t start = rowidx.size();
                for (const auto &o : ord) {
                    if (rowidx.size() > start && rowidx.back() == o.first) val.back() += va[(size_t)o.second];
                    else { rowidx.push_back(o.first); val.push_back(va[(size_t)o.second]); }
                }
            }
            colptr[(size_t)j + 1] = (int64_t)val.size();
        }
    }

    SparseMatrixD transpose() const
    {
        SparseMatrixD t;
        t.nrows = ncols; t.ncols = nrows;
        t.colptr.assign((size_t)nrows + 1, 0);
        for (int32_t r : rowidx) t.colptr[(size_t)r + 1]++;
        for (int64_t i = 0; i < nrows; ++i) t.colptr[(size_t)i + 1] += t.colptr[(size_t)i];
        t.rowidx.resize(val.size()); t.val.resize(val.size());
        std::vector<int64_t> pos(t.colptr.begin(), t.colptr.end() - 1);
        for (int64_t j = 0; j < ncols; ++j)
            for (int64_t p = colptr[(size_t)j]; p < colptr[(size_t)j + 1]; ++p) {
                const int64_t q = pos[(size_t)rowidx[(size_t)p]]++;
                t.rowidx[(size_t)q] = (int32_t)j; t.val[(size_t)q] = val[(size_t)p];
            }
        return t;
    }

    // Eigen conservativeResize to dims that are >= the current ones (c++/sample.cpp:119-122)
    void conservativeResize(int64_t nr, int64_t nc)
    {
        if (nr < nrows || nc < ncols) throw std::runtime_error("conservativeResize: shrinking is not supported");
        nrows = nr;
        colptr.resize((size_t)nc + 1, colptr.back());
        ncols = nc;
    }
};

}  // namespace bpmf_host
