// fast_kernels.cu — the fused per-item conditional update for num_latent == 32 on sm_100a:
// one warp per item (c++/sample.cpp:263-336 + 248-258 in a single kernel).
//
//   stage 1  Gram + rhs   for each group of 4 ratings every lane gathers 32 B of one rating's latent
//                         vector (lane = 4*q + r reads dims 4q..4q+3 of rating r), and the warp issues
//                         10 fp64 tensor-core DMMAs (mma.sync.m8n8k4) for the upper triangle of 4x4
//                         blocks of 8 latent dims ("block a" = dims {4q'+a}); the same register is the
//                         A fragment of block a and the B fragment of block a, so a gathered value is
//                         loaded once and never moved. rr is 4 DFMAs per group.
//   stage 2  MM = LambdaF + alpha*G, transposed through shared memory to "lane j owns column j"
//   stage 3  Cholesky in registers (right-looking on the full symmetric matrix: lane j scales its own
//            entry, column k of L is broadcast through shared memory)
//   stage 4  forward solve, + K normals (Philox4x32-10, polar method, warp-ballot numbering),
//            backward solve, coalesced 256-byte store (and stores into every peer replica)
//
// Roofline: per rating 256 B gathered and 2.5 DMMA = 640 fp64 FMA; the DMMA pipe
// (measured 37 TFLOP/s) and HBM (6.5 TB/s) co-limit at ~7.2 TB/s-equivalent.
#include "common.cuh"
#include "rng.cuh"

namespace bpmf {

constexpr int FW = 4;          // warps per CTA
constexpr int LS = 33;         // padded stride (doubles) of the per-warp 32x32 tile
constexpr int WARP_SMEM = 32 * LS + 64;   // tile + z[32] + b[32]

struct FastArgs {
    int from, to;
    uint32_t iter;
    double alpha, mean_rating;
    const int64_t *colptr;
    const int32_t *rowidx;
    const double *val;
    const double *other;
    double *items;
    int npeers;
    double *const *peers;
    const double *mu, *LambdaF;
    unsigned int *work_counter;
    unsigned long long *err;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(FW * 32, 4) items_dmma32_kernel(FastArgs p)
{
    extern __shared__ double sm[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *sLF = sm;                      // LambdaF, column stride LS
    double *srr0 = sLF + 32 * LS;          // LambdaF * mu
    double *wt = srr0 + 32 + warp * WARP_SMEM;
    double *wz = wt + 32 * LS;
    double *wb = wz + 32;

    for (int e = tid; e < 1024; e += FW * 32) sLF[(e & 31) + (e >> 5) * LS] = p.LambdaF[e];
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = 0; j < 32; ++j) s += sLF[tid + j * LS] * p.mu[j];   // rr = LambdaF * hp.mu (sample.cpp:285)
        srr0[tid] = s;
    }
    __syncthreads();

    const int q = lane >> 2, r = lane & 3;
    int next = 0;
    if (lane == 0) next = p.from + (int)atomicAdd(p.work_counter, 1u);
    for (;;) {
        const int idx = __shfl_sync(FULL, next, 0);
        if (idx >= p.to) break;
        if (lane == 0) next = p.from + (int)atomicAdd(p.work_counter, 1u);   // claim the next item early

        const int64_t ps = p.colptr[idx], pe = p.colptr[idx + 1];
        // ---- stage 1: G (upper triangle of blocks) and rr --------------------------------------------
        double c[10][2];
#pragma unroll
        for (int t = 0; t < 10; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
        double rrp[4] = {0.0, 0.0, 0.0, 0.0};
        int32_t myidx = -1;
        double myw = 0.0;
        if (ps + lane < pe) {
            myidx = p.rowidx[ps + lane];
            myw = (p.val[ps + lane] - p.mean_rating) * p.alpha;
        }
        for (int64_t t0 = ps; t0 < pe; t0 += 32) {
            const int cur_idx = myidx;
            const double cur_w = myw;
            // prefetch the next chunk's indices and weights
            myidx = -1; myw = 0.0;
            if (t0 + 32 + lane < pe) {
                myidx = p.rowidx[t0 + 32 + lane];
                myw = (p.val[t0 + 32 + lane] - p.mean_rating) * p.alpha;
            }
            const int left = (int)min((int64_t)32, pe - t0);
            const int ng = (left + 3) >> 2;
            // issue every gather of the chunk before the first DMMA (8 x 32 B per lane in flight)
            double y[8][4];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int j = __shfl_sync(FULL, cur_idx, 4 * g + r);
                y[g][0] = y[g][1] = y[g][2] = y[g][3] = 0.0;
                if (j >= 0) {
                    const double2 *src = reinterpret_cast<const double2 *>(p.other + (size_t)j * 32 + 4 * q);
                    const double2 lo = __ldg(src), hi = __ldg(src + 1);
                    y[g][0] = lo.x; y[g][1] = lo.y; y[g][2] = hi.x; y[g][3] = hi.y;
                }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (g < ng) {   // warp-uniform
                    int t = 0;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = a; b < 4; ++b) { dmma884(c[t][0], c[t][1], y[g][a], y[g][b]); ++t; }
                    const double wg = __shfl_sync(FULL, cur_w, 4 * g + r);
#pragma unroll
                    for (int a = 0; a < 4; ++a) rrp[a] = fma(y[g][a], wg, rrp[a]);
                }
            }
        }
        // rr: sum the 4 ratings of a group held by the quad
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 1);
            rrp[a] += __shfl_xor_sync(FULL, rrp[a], 2);
        }
        __syncwarp();   // previous item's readers of wt/wb/wz are done
        if (r == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) wb[4 * q + a] = srr0[4 * q + a] + rrp[a];
        }
        // ---- stage 2: scatter G (both triangles) into the warp tile --------------------------------
        {
            int t = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a; b < 4; ++b) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int row = 4 * q + a, col = 4 * (2 * r + i) + b;
                        wt[row * LS + col] = c[t][i];
                        if (a != b) wt[col * LS + row] = c[t][i];
                    }
                    ++t;
                }
        }
        // the K normals of this item: rng_set_pos((idx+1)*K*(iter+1)) (sample.cpp:266)
        warp_randn((uint32_t)(((long long)idx + 1) * 32ll * ((long long)p.iter + 1)), 32, wz);
        __syncwarp();
        // lane j takes column j of MM = LambdaF + alpha * G (sample.cpp:297-298); only the lower triangle of
        // LambdaF is referenced, like the LLT of the reference
        double A[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const double lf = (i >= lane) ? sLF[i + lane * LS] : sLF[lane + i * LS];
            A[i] = fma(p.alpha, wt[i * LS + lane], lf);
        }
        double b = wb[lane];
        const double z = wz[lane];
        __syncwarp();
        // ---- stage 3: Cholesky (sample.cpp:306) ------------------------------------------------------
        double invd = 0.0;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const double d = __shfl_sync(FULL, A[k], k);
            if (!(d > 0.0)) ok = false;   // Eigen: pivot <= 0 -> NumericalIssue ("Cholesky failed")
            const double inv = 1.0 / sqrt(d);
            const double l = A[k] * inv;  // lane j > k: L(j,k); lane k: sqrt(d)
            A[k] = l;
            if (lane == k) invd = inv;
            wt[k * LS + lane] = l;        // column k of L
            __syncwarp();
#pragma unroll
            for (int i = k + 1; i < 32; ++i) A[i] = fma(-wt[k * LS + i], l, A[i]);
        }
        if (!ok) {
            if (lane == 0) atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
            continue;
        }
        // ---- stage 4: L y = rr ; y += z ; L^T x = y (sample.cpp:321-323) -----------------------------
        double yv = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const double yk = __shfl_sync(FULL, b * invd, k);
            if (lane == k) yv = yk;
            b = fma(-A[k], yk, b);        // meaningful for lanes > k
        }
        yv += z;
        double xv = 0.0;
#pragma unroll
        for (int k = 31; k >= 0; --k) {
            const double xk = __shfl_sync(FULL, yv * invd, k);
            if (lane == k) xv = xk;
            yv = fma(-wt[lane * LS + k], xk, yv);   // L(k,lane), meaningful for lanes < k
        }
        // items().col(idx) = rr (sample.cpp:324); push to the peer replicas (replaces send_item, :370)
        p.items[(size_t)idx * 32 + lane] = xv;
        for (int pr = 0; pr < p.npeers; ++pr) {
            double *dst = p.peers[pr];
            if (dst && dst != p.items) dst[(size_t)idx * 32 + lane] = xv;
        }
    }
}

cudaError_t launch_items_dmma32(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    FastArgs p;
    p.from = s.from; p.to = s.to; p.iter = iter; p.alpha = alpha; p.mean_rating = s.mean_rating;
    p.colptr = s.colptr; p.rowidx = s.rowidx; p.val = s.val;
    p.other = o.items; p.items = s.items;
    p.npeers = s.npeers; p.peers = s.peers_dev;
    p.mu = s.hp.mu; p.LambdaF = s.hp.LambdaF;
    p.work_counter = s.work_counter; p.err = c->d_err;
    cudaError_t e = cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned int), c->stream);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(double) * (32 * LS + 32 + FW * WARP_SMEM);
    e = cudaFuncSetAttribute(items_dmma32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, items_dmma32_kernel, FW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long n = (long long)s.to - s.from;
    long long grid = (long long)c->sm_count * per_sm;
    const long long need = (n + FW - 1) / FW;
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    items_dmma32_kernel<<<(unsigned)grid, FW * 32, smem, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace bpmf
