// io.cpp — see io.h. Whole-file reads into memory (through zlib when the name ends in .gz) and a
// cursor-based parser; no iostream token extraction, so a 100M-entry .sdm or .mtx loads at disk speed.
#include "io.h"

#include <cctype>
#include <charconv>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace bpmf_host {

namespace {

[[noreturn]] void fail(const std::string &msg) { throw std::runtime_error(msg); }

std::string upper(std::string s)
{
    for (char &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}

// ---- raw bytes in / out -------------------------------------------------------------------------
std::string slurp(const std::string &filename, bool compressed)
{
    std::string buf;
    if (compressed) {
        gzFile f = gzopen(filename.c_str(), "rb");
        if (!f) fail("File " + filename + " does not exist or cannot be opened");
        gzbuffer(f, 1 << 20);
        char tmp[1 << 16];
        int n;
        while ((n = gzread(f, tmp, sizeof tmp)) > 0) buf.append(tmp, (size_t)n);
        const bool bad = n < 0;
        gzclose(f);
        if (bad) fail("Error while decompressing " + filename);
    } else {
        FILE *f = fopen(filename.c_str(), "rb");
        if (!f) fail("File " + filename + " does not exist or cannot be opened");
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        const size_t got = buf.empty() ? 0 : fread(&buf[0], 1, buf.size(), f);
        fclose(f);
        if (got != buf.size()) fail("Short read on " + filename);
    }
    return buf;
}

void spill(const std::string &filename, bool compressed, const std::string &bytes)
{
    if (compressed) {
        gzFile f = gzopen(filename.c_str(), "wb");
        if (!f) fail("Cannot open " + filename + " for writing");
        size_t off = 0;
        while (off < bytes.size()) {
            const unsigned chunk = (unsigned)std::min<size_t>(bytes.size() - off, 1u << 30);
            if (gzwrite(f, bytes.data() + off, chunk) != (int)chunk) { gzclose(f); fail("Error while writing " + filename); }
            off += chunk;
        }
        gzclose(f);
    } else {
        FILE *f = fopen(filename.c_str(), "wb");
        if (!f) fail("Cannot open " + filename + " for writing");
        const size_t put = bytes.empty() ? 0 : fwrite(bytes.data(), 1, bytes.size(), f);
        fclose(f);
        if (put != bytes.size()) fail("Error while writing " + filename);
    }
}

// ---- binary cursor --------------------------------------------------------------------------------
struct BinCursor {
    const std::string &b;
    size_t off = 0;
    explicit BinCursor(const std::string &s) : b(s) {}
    void take(void *dst, size_t n)
    {
        if (n > b.size() - off) fail("Unexpected end of binary matrix file");   // (off <= size always; no wrap-around)
        memcpy(dst, b.data() + off, n);
        off += n;
    }
    uint64_t u64() { uint64_t v; take(&v, 8); return v; }
};

template <class Sink>
void read_sparse_bin(const std::string &bytes, bool with_values, Sink &X)
{
    BinCursor c(bytes);
    const uint64_t nrow = c.u64(), ncol = c.u64(), nnz = c.u64();
    // sizes come from the file: bound them by what the file can hold BEFORE multiplying (a crafted header must not wrap)
    if (nnz > (bytes.size() - 24) / (with_values ? 16 : 8)) fail("Invalid number of values");
    if (nrow > 0x7fffffffull || ncol > 0x7fffffffull) fail("Matrix dimensions exceed the 32-bit index range");
    std::vector<uint32_t> rows(nnz), cols(nnz);
    c.take(rows.data(), nnz * 4);
    c.take(cols.data(), nnz * 4);
    std::vector<double> vals;
    if (with_values) { vals.resize(nnz); c.take(vals.data(), nnz * 8); }
    std::vector<Triplet> t(nnz);
    for (uint64_t i = 0; i < nnz; ++i) {   // indices are 1-based on disk (io.cpp:268-272)
        t[i].row = (int32_t)(rows[i] - 1); t[i].col = (int32_t)(cols[i] - 1);
        t[i].val = with_values ? vals[i] : 1.0;
    }
    X.from_triplets((int64_t)nrow, (int64_t)ncol, t);
    if (with_values && X.nonZeros() != (int64_t)nnz) fail("Invalid number of values");   // io.cpp:284-287
}

void read_dense_bin(const std::string &bytes, DenseMatrixD &X)
{
    BinCursor c(bytes);
    const uint64_t nrow = c.u64(), ncol = c.u64();
    const uint64_t cap = (bytes.size() - 16) / 8;           // doubles the file holds after its header
    if ((ncol != 0 && nrow > cap / ncol) || nrow * ncol > cap) fail("Invalid dense matrix size");   // also refuses a wrapping product
    X.resize((int64_t)nrow, (int64_t)ncol);
    c.take(X.data(), (size_t)(nrow * ncol) * 8);
}

// ---- text cursor ----------------------------------------------------------------------------------
struct TextCursor {
    const char *p, *end;
    explicit TextCursor(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
    TextCursor(const char *b, const char *e) : p(b), end(e) {}
    bool eof() const { return p >= end; }
    std::string line()
    {
        const char *q = (const char *)memchr(p, '\n', (size_t)(end - p));
        std::string s(p, q ? q : end);
        p = q ? q + 1 : end;
        if (!s.empty() && s.back() == '\r') s.pop_back();
        return s;
    }
    // comments ('%' lines) and empty lines may sit before the size line and before every entry (io.cpp:377-378,393-395)
    void skip_comments()
    {
        for (;;) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
            if (p < end && *p == '\n') { ++p; continue; }
            if (p < end && *p == '%') { const char *q = (const char *)memchr(p, '\n', (size_t)(end - p)); p = q ? q + 1 : end; continue; }
            return;
        }
    }
    void skip_ws() { while (p < end && std::isspace((unsigned char)*p)) ++p; }
    bool u64(uint64_t &v)
    {
        skip_ws();
        if (p >= end || !std::isdigit((unsigned char)*p)) return false;
        uint64_t x = 0;
        while (p < end && std::isdigit((unsigned char)*p)) x = x * 10 + (uint64_t)(*p++ - '0');
        v = x;
        return true;
    }
    bool f64(double &v)
    {
        skip_ws();
        if (p >= end) return false;
        const char *q = (*p == '+') ? p + 1 : p;               // from_chars does not take a leading '+'
        const std::from_chars_result r = std::from_chars(q, end, v);
        if (r.ec == std::errc() ) { p = r.ptr; return true; }
        // "inf" / "nan" spellings and out-of-range values: fall back to strtod on a bounded copy
        char tmp[64];
        const size_t n = std::min<size_t>(sizeof tmp - 1, (size_t)(end - p));
        memcpy(tmp, p, n);
        tmp[n] = 0;
        char *e = nullptr;
        v = strtod(tmp, &e);
        if (e == tmp) return false;
        p += (e - tmp);
        return true;
    }
};

struct MMHeader { std::string object, format, field, symmetry; };

MMHeader mm_header(TextCursor &c)
{
    // "%%MatrixMarket" followed by a blank, then four case-insensitive tokens (io.cpp:417-448)
    const std::string first = c.line();
    if (first.size() < 15 || first.compare(0, 14, "%%MatrixMarket") != 0 || !(first[14] == ' ' || first[14] == '\t'))
        fail("Cannot read MatrixMarket from input stream: the first 15 characters must be '%%MatrixMarket' followed by at "
             "least one blank\nGot: " + first.substr(0, 14));
    MMHeader h;
    std::string *dst[4] = {&h.object, &h.format, &h.field, &h.symmetry};
    size_t i = 14;
    for (int k = 0; k < 4; ++k) {
        while (i < first.size() && std::isspace((unsigned char)first[i])) ++i;
        size_t j = i;
        while (j < first.size() && !std::isspace((unsigned char)first[j])) ++j;
        *dst[k] = upper(first.substr(i, j - i));
        i = j;
    }
    if (h.object != "MATRIX") fail("Invalid MartrixMarket object type: expected 'matrix' but got '" + h.object + "'");
    if (h.symmetry != "GENERAL") fail("Invalid MatrixMarket symmetry type: only 'general' symmetry type is supported");
    return h;
}

template <class Sink>
void read_mm_sparse(const std::string &bytes, Sink &X)
{
    TextCursor c(bytes);
    const MMHeader h = mm_header(c);
    if (h.field != "REAL" && h.field != "PATTERN")
        fail("Invalid MatrixMarket field type: only 'real' and 'pattern' field types are supported");
    c.skip_comments();
    if (h.format != "COORDINATE") fail("Cannot read a dense matrix as a sparse matrix");
    uint64_t nrows, ncols, nnz;
    if (!c.u64(nrows) || !c.u64(ncols) || !c.u64(nnz)) fail("Could not get 'rows', 'cols', 'nnz' values for coordinate matrix format");
    const bool pattern = h.field == "PATTERN";
    // The entry lines are parsed in parallel: the rest of the buffer is cut at line ends into one piece per thread, every
    // piece is parsed into its own list, and the lists are concatenated in file order (duplicates are summed in input
    // order by from_triplets, as Eigen's setFromTriplets does).
    const char *b = c.p, *e = c.end;
    int nt = 1;
#ifdef _OPENMP
    nt = std::max(1, std::min(omp_get_max_threads(), (int)((e - b) / (1 << 20)) + 1));
#endif
    std::vector<const char *> cut((size_t)nt + 1, e);
    cut[0] = b;
    for (int i = 1; i < nt; ++i) {
        const char *q = b + (size_t)(e - b) * (size_t)i / (size_t)nt;
        q = (const char *)memchr(q, '\n', (size_t)(e - q));
        cut[(size_t)i] = q ? q + 1 : e;
    }
    std::vector<std::vector<Triplet>> part((size_t)nt);
    std::vector<int> bad((size_t)nt, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nt)
#endif
    for (int i = 0; i < nt; ++i) {
        TextCursor pc(cut[(size_t)i], cut[(size_t)i + 1]);
        std::vector<Triplet> &out = part[(size_t)i];
        out.reserve((size_t)(nnz / (uint64_t)nt) + 16);
        for (;;) {
            pc.skip_comments();
            if (pc.eof()) break;
            uint64_t r, col;
            double v = 1.0;
            if (!pc.u64(r) || !pc.u64(col) || (!pattern && !pc.f64(v))) { bad[(size_t)i] = 1; break; }
            Triplet tr;
            tr.row = (int32_t)(r - 1); tr.col = (int32_t)(col - 1); tr.val = v;
            out.push_back(tr);
        }
    }
    std::vector<Triplet> t;
    t.reserve((size_t)nnz);
    for (int i = 0; i < nt && t.size() < (size_t)nnz; ++i) {
        const size_t take = std::min(part[(size_t)i].size(), (size_t)nnz - t.size());
        t.insert(t.end(), part[(size_t)i].begin(), part[(size_t)i].begin() + (std::ptrdiff_t)take);
        // a malformed line only matters if it comes before the nnz-th entry (the reference stops reading there)
        if (bad[(size_t)i] && t.size() < (size_t)nnz) fail("Could not parse an entry line for coordinate matrix format");
        std::vector<Triplet>().swap(part[(size_t)i]);
    }
    if (t.size() < (size_t)nnz) fail("Could not parse an entry line for coordinate matrix format");
    X.from_triplets((int64_t)nrows, (int64_t)ncols, t);
}

void read_mm_dense(const std::string &bytes, DenseMatrixD &X)
{
    TextCursor c(bytes);
    const MMHeader h = mm_header(c);
    if (h.field != "REAL") fail("Invalid MatrixMarket field type: only 'real' field type is supported");
    c.skip_comments();
    if (h.format != "ARRAY") fail("Cannot read a sparse matrix as a dense matrix");
    uint64_t nrows, ncols;
    if (!c.u64(nrows) || !c.u64(ncols)) fail("Could not get 'rows', 'cols' values for array matrix format");
    X.resize((int64_t)nrows, (int64_t)ncols);
    for (uint64_t col = 0; col < ncols; ++col)
        for (uint64_t row = 0; row < nrows; ++row) {
            c.skip_comments();
            double v;
            if (!c.f64(v)) fail("Could not parse an entry line for array matrix format");
            X((int64_t)row, (int64_t)col) = v;
        }
}

void read_csv_dense(const std::string &bytes, DenseMatrixD &X)
{
    TextCursor c(bytes);
    const uint64_t nrow = strtoull(c.line().c_str(), nullptr, 10);
    const uint64_t ncol = strtoull(c.line().c_str(), nullptr, 10);
    X.resize((int64_t)nrow, (int64_t)ncol);
    uint64_t row = 0, col = 0;
    while (!c.eof() && row < nrow) {
        const std::string l = c.line();
        col = 0;
        size_t i = 0;
        while (i <= l.size() && col < ncol) {
            size_t j = l.find(',', i);
            if (j == std::string::npos) j = l.size();
            X((int64_t)row, (int64_t)col++) = strtod(l.substr(i, j - i).c_str(), nullptr);
            i = j + 1;
        }
        ++row;
    }
    if (row != nrow) fail("invalid number of rows");
    if (col != ncol) fail("invalid number of columns");
}

// ---- writers ----------------------------------------------------------------------------------------
void put(std::string &s, const void *p, size_t n) { s.append((const char *)p, n); }
void put_g(std::string &s, double v)   // std::ostream << double at default precision == %g
{
    char b[40];
    s.append(b, (size_t)snprintf(b, sizeof b, "%g", v));
}
void put_u(std::string &s, uint64_t v)
{
    char b[24];
    s.append(b, (size_t)snprintf(b, sizeof b, "%llu", (unsigned long long)v));
}

std::string dense_bytes(MatrixType mt, const double *x, int64_t nrows, int64_t ncols)
{
    std::string s;
    const uint64_t nr = (uint64_t)nrows, nc = (uint64_t)ncols;
    switch (mt.type) {
    case MatrixType::ddm:   // io.cpp:607-615
        put(s, &nr, 8); put(s, &nc, 8); put(s, x, (size_t)(nr * nc) * 8);
        break;
    case MatrixType::mtx:   // io.cpp:686-702
        s += "%%MatrixMarket MATRIX ARRAY REAL GENERAL\n";
        put_u(s, nr); s += ' '; put_u(s, nc); s += '\n';
        for (uint64_t i = 0; i < nr * nc; ++i) { put_g(s, x[i]); s += '\n'; }
        break;
    case MatrixType::csv:   // io.cpp:617-624: Eigen IOFormat(6, DontAlignCols, ",", "\n")
        put_u(s, nr); s += '\n'; put_u(s, nc); s += '\n';
        for (uint64_t r = 0; r < nr; ++r) {
            for (uint64_t c = 0; c < nc; ++c) { if (c) s += ','; put_g(s, x[r + c * nr]); }
            s += '\n';
        }
        break;
    default:
        fail("Invalid matrix type for a dense matrix");
    }
    return s;
}

}  // namespace

MatrixType ExtensionToMatrixType(const std::string &fname)
{
    size_t dot = fname.find_last_of('.');
    if (dot == std::string::npos) fail("Extension is not specified in " + fname);
    std::string ext = fname.substr(dot);
    bool compressed = false;
    if (ext == ".gz") {
        compressed = true;
        const size_t dot2 = dot ? fname.find_last_of('.', dot - 1) : std::string::npos;
        if (dot2 == std::string::npos) fail("Extension is not specified in " + fname);
        ext = fname.substr(dot2, dot - dot2);
    }
    if (ext == ".sdm") return {MatrixType::sdm, compressed};
    if (ext == ".sbm") return {MatrixType::sbm, compressed};
    if (ext == ".mtx" || ext == ".mm") return {MatrixType::mtx, compressed};
    if (ext == ".csv") return {MatrixType::csv, compressed};
    if (ext == ".ddm") return {MatrixType::ddm, compressed};
    fail("Unknown file type: " + ext + " specified in " + fname);
}

void read_matrix(const std::string &filename, SparseMatrixD &X)
{
    const MatrixType mt = ExtensionToMatrixType(filename);
    const std::string bytes = slurp(filename, mt.compressed);
    switch (mt.type) {
    case MatrixType::sdm: read_sparse_bin(bytes, true, X); break;
    case MatrixType::sbm: read_sparse_bin(bytes, false, X); break;
    case MatrixType::mtx: read_mm_sparse(bytes, X); break;
    default: fail("Invalid matrix type: " + filename + " is a dense format, a sparse matrix was asked for");
    }
}

void read_matrix(const std::string &filename, TripletList &X)
{
    const MatrixType mt = ExtensionToMatrixType(filename);
    const std::string bytes = slurp(filename, mt.compressed);
    switch (mt.type) {
    case MatrixType::sdm: read_sparse_bin(bytes, true, X); X.refuse_duplicates = true; break;   // refused by the caller once they are summed
    case MatrixType::sbm: read_sparse_bin(bytes, false, X); break;
    case MatrixType::mtx: read_mm_sparse(bytes, X); break;
    default: fail("Invalid matrix type: " + filename + " is a dense format, a sparse matrix was asked for");
    }
}

void read_matrix(const std::string &filename, DenseMatrixD &X)
{
    const MatrixType mt = ExtensionToMatrixType(filename);
    const std::string bytes = slurp(filename, mt.compressed);
    switch (mt.type) {
    case MatrixType::ddm: read_dense_bin(bytes, X); break;
    case MatrixType::csv: read_csv_dense(bytes, X); break;
    case MatrixType::mtx: read_mm_dense(bytes, X); break;
    default: fail("Invalid matrix type: " + filename + " is a sparse format, a dense matrix was asked for");
    }
}

void write_matrix(const std::string &filename, const SparseMatrixD &X)
{
    const MatrixType mt = ExtensionToMatrixType(filename);
    std::string s;
    const uint64_t nr = (uint64_t)X.nrows, nc = (uint64_t)X.ncols;
    if (mt.type == MatrixType::sdm || mt.type == MatrixType::sbm) {   // io.cpp:626-682: entries in column order, 1-based
        const bool with_values = mt.type == MatrixType::sdm;
        std::vector<uint32_t> rows, cols;
        std::vector<double> vals;
        for (int64_t j = 0; j < X.ncols; ++j)
            for (int64_t p = X.colptr[(size_t)j]; p < X.colptr[(size_t)j + 1]; ++p) {
                if (!with_values && !(X.val[(size_t)p] > 0)) continue;
                rows.push_back((uint32_t)X.rowidx[(size_t)p] + 1); cols.push_back((uint32_t)j + 1);
                if (with_values) vals.push_back(X.val[(size_t)p]);
            }
        const uint64_t nnz = rows.size();
        put(s, &nr, 8); put(s, &nc, 8); put(s, &nnz, 8);
        put(s, rows.data(), nnz * 4); put(s, cols.data(), nnz * 4);
        if (with_values) put(s, vals.data(), nnz * 8);
    } else if (mt.type == MatrixType::mtx) {                          // io.cpp:704-719
        s += "%%MatrixMarket MATRIX COORDINATE REAL GENERAL\n";
        put_u(s, nr); s += ' '; put_u(s, nc); s += ' '; put_u(s, (uint64_t)X.nonZeros()); s += '\n';
        for (int64_t j = 0; j < X.ncols; ++j)
            for (int64_t p = X.colptr[(size_t)j]; p < X.colptr[(size_t)j + 1]; ++p) {
                put_u(s, (uint64_t)X.rowidx[(size_t)p] + 1); s += ' '; put_u(s, (uint64_t)j + 1); s += ' ';
                put_g(s, X.val[(size_t)p]); s += '\n';
            }
    } else {
        fail("Invalid matrix type for a sparse matrix");
    }
    spill(filename, mt.compressed, s);
}

void write_matrix(const std::string &filename, const double *colmajor, int64_t nrows, int64_t ncols)
{
    const MatrixType mt = ExtensionToMatrixType(filename);
    spill(filename, mt.compressed, dense_bytes(mt, colmajor, nrows, ncols));
}

void write_matrix(const std::string &filename, const DenseMatrixD &X) { write_matrix(filename, X.data(), X.nrows, X.ncols); }

}  // namespace bpmf_host
