// Device-side build of a side's compressed matrix from a coordinate list.
//
// Replaces, for inputs that are already in memory as (row, col, value) triplets, what the reference does on the host with
// Eigen: setFromTriplets (c++/io.cpp:282,521: column-major, inner indices ascending, duplicates summed) and the transpose
// that gives the other factor its matrix (c++/sample.cpp:133-134). Both orientations come from the same triplet arrays:
//
//   key   = major << 32 | minor, stable LSD radix sort (cub::DeviceRadixSort, the one library call; load path, not the
//           sweep) carrying the entry's input position
//   heads = first entry of every run of equal keys; an exclusive scan of the head flags numbers the output entries
//   emit  = one thread per head walks its run in INPUT order and adds the values sequentially, so duplicates are summed
//           exactly like the host build (matrix.h from_triplets) sums them
//   ptr   = colptr[j] = output position of the first sorted entry whose key is >= j << 32 (binary search per j)
//
// The result is bit-identical to the host build (tests/test_gpu_parity.py::test_device_build_*).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace bpmf {
namespace {

__global__ void make_keys_kernel(int64_t n, const int32_t *__restrict__ major, const int32_t *__restrict__ minor, int num_major,
                                 int num_minor, unsigned long long *__restrict__ keys, uint32_t *__restrict__ perm, int *__restrict__ bad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = major[i], b = minor[i];
        if (a < 0 || a >= num_major || b < 0 || b >= num_minor) *bad = 1;
        keys[i] = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
        perm[i] = (uint32_t)i;
    }
}

struct HeadFlag {
    const unsigned long long *keys;
    __host__ __device__ long long operator()(long long i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1ll : 0ll; }
};

__global__ void emit_kernel(int64_t n, const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ perm,
                            const long long *__restrict__ pos, const double *__restrict__ val_in, int32_t *__restrict__ idx_out,
                            int32_t *__restrict__ major_out, double *__restrict__ val_out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (i > 0 && keys[i - 1] == k) continue;           // not the head of its run
        double acc = val_in[perm[i]];
        for (int64_t j = i + 1; j < n && keys[j] == k; ++j) acc += val_in[perm[j]];   // duplicates, in input order
        const long long r = pos[i];
        idx_out[r] = (int32_t)(uint32_t)(k & 0xffffffffull);
        if (major_out) major_out[r] = (int32_t)(k >> 32);
        val_out[r] = acc;
    }
}

__global__ void colptr_kernel(int num_major, int64_t n, long long n_out, const unsigned long long *__restrict__ keys,
                              const long long *__restrict__ pos, int64_t *__restrict__ colptr)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j <= num_major; j += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long want = (unsigned long long)j << 32;
        int64_t lo = 0, hi = n;                            // first sorted entry with key >= want
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < want) lo = mid + 1; else hi = mid;
        }
        colptr[j] = (lo == n) ? n_out : pos[lo];           // that entry starts a run, so pos[] is its output position
    }
}

struct Scratch {
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *p0 = nullptr, *p1 = nullptr;
    long long *pos = nullptr;
    void *tmp = nullptr;
    int *bad = nullptr;
    ~Scratch() { cudaFree(k0); cudaFree(k1); cudaFree(p0); cudaFree(p1); cudaFree(pos); cudaFree(tmp); cudaFree(bad); }
};

int bits_for(int n)
{
    int b = 1;
    while (b < 32 && (1ll << b) < (long long)n) ++b;
    return b;
}

}  // namespace

// d_major / d_minor / d_val: the triplets on the device. Outputs are cudaMalloc'ed here (idx / val with 32 entries of
// zero padding, as load_side allocates them); *bad_index is set when an index is out of range.
cudaError_t build_compressed(bpmf_gpu_ctx *c, int64_t n, int num_major, int num_minor, const int32_t *d_major, const int32_t *d_minor,
                             const double *d_val, int64_t **colptr_out, int32_t **idx_out, int32_t **major_out, double **val_out,
                             int64_t *nnz_out, bool *bad_index)
{
#define BC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)
    *colptr_out = nullptr; *idx_out = nullptr; *val_out = nullptr; *nnz_out = 0; *bad_index = false;
    if (major_out) *major_out = nullptr;
    cudaStream_t st = c->stream;
    const int grid = c->sm_count * 8;
    BC(cudaMalloc(colptr_out, sizeof(int64_t) * ((size_t)num_major + 1)));
    if (n == 0) {
        BC(cudaMemsetAsync(*colptr_out, 0, sizeof(int64_t) * ((size_t)num_major + 1), st));
        BC(cudaMalloc(idx_out, sizeof(int32_t) * 32)); BC(cudaMemsetAsync(*idx_out, 0, sizeof(int32_t) * 32, st));
        BC(cudaMalloc(val_out, sizeof(double) * 32)); BC(cudaMemsetAsync(*val_out, 0, sizeof(double) * 32, st));
        if (major_out) { BC(cudaMalloc(major_out, sizeof(int32_t) * 32)); BC(cudaMemsetAsync(*major_out, 0, sizeof(int32_t) * 32, st)); }
        return cudaStreamSynchronize(st);
    }
    Scratch s;
    BC(cudaMalloc(&s.k0, sizeof(unsigned long long) * n)); BC(cudaMalloc(&s.k1, sizeof(unsigned long long) * n));
    BC(cudaMalloc(&s.p0, sizeof(uint32_t) * n)); BC(cudaMalloc(&s.p1, sizeof(uint32_t) * n));
    BC(cudaMalloc(&s.pos, sizeof(long long) * n));
    BC(cudaMalloc(&s.bad, sizeof(int)));
    BC(cudaMemsetAsync(s.bad, 0, sizeof(int), st));
    make_keys_kernel<<<grid, 256, 0, st>>>(n, d_major, d_minor, num_major, num_minor, s.k0, s.p0, s.bad);
    BC(cudaGetLastError());
    int bad = 0;
    BC(cudaMemcpyAsync(&bad, s.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    BC(cudaStreamSynchronize(st));
    if (bad) { *bad_index = true; return cudaSuccess; }
    // stable sort by (major, minor): only the bits that can be set take part
    const int end_bit = 32 + bits_for(num_major);
    size_t tmp_bytes = 0, scan_bytes = 0;
    BC(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s.k0, s.k1, s.p0, s.p1, n, 0, end_bit, st));
    auto heads = thrust::make_transform_iterator(thrust::counting_iterator<long long>(0), HeadFlag{s.k1});
    BC(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, heads, s.pos, n, st));
    BC(cudaMalloc(&s.tmp, tmp_bytes > scan_bytes ? tmp_bytes : scan_bytes));
    BC(cub::DeviceRadixSort::SortPairs(s.tmp, tmp_bytes, s.k0, s.k1, s.p0, s.p1, n, 0, end_bit, st));
    BC(cub::DeviceScan::ExclusiveSum(s.tmp, scan_bytes, heads, s.pos, n, st));
    long long last_pos = 0;
    BC(cudaMemcpyAsync(&last_pos, s.pos + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
    BC(cudaStreamSynchronize(st));
    // pos[] counts the heads BEFORE an entry: the number of runs is pos[n-1] plus one if the last entry is itself a head
    unsigned long long tail_keys[2] = {0, 1};
    if (n > 1) BC(cudaMemcpy(tail_keys, s.k1 + (n - 2), 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const long long runs = last_pos + (tail_keys[0] != tail_keys[1] ? 1 : 0);
    BC(cudaMalloc(idx_out, sizeof(int32_t) * (size_t)(runs + 32)));
    BC(cudaMalloc(val_out, sizeof(double) * (size_t)(runs + 32)));
    BC(cudaMemsetAsync(*idx_out, 0, sizeof(int32_t) * (size_t)(runs + 32), st));
    BC(cudaMemsetAsync(*val_out, 0, sizeof(double) * (size_t)(runs + 32), st));
    if (major_out) {
        BC(cudaMalloc(major_out, sizeof(int32_t) * (size_t)(runs + 32)));
        BC(cudaMemsetAsync(*major_out, 0, sizeof(int32_t) * (size_t)(runs + 32), st));
    }
    emit_kernel<<<grid, 256, 0, st>>>(n, s.k1, s.p1, s.pos, d_val, *idx_out, major_out ? *major_out : nullptr, *val_out);
    BC(cudaGetLastError());
    colptr_kernel<<<grid, 256, 0, st>>>(num_major, n, runs, s.k1, s.pos, *colptr_out);
    BC(cudaGetLastError());
    BC(cudaStreamSynchronize(st));
    *nnz_out = runs;
    c->launches += 3;
    return cudaSuccess;
#undef BC
}

}  // namespace bpmf
