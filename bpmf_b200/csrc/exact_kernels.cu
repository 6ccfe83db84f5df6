// exact_kernels.cu — reference-order device kernels. COMPILED WITH -fmad=false so that every
// floating-point operation rounds exactly like the reference's default x86-64 build (no FMA
// contraction); summation orders are the ones the CPU checker used by tests/ states, which follow
// c++/sample.cpp and c++/mvnormal.cpp. Contains:
//   hyper_kernel        CondNormalWishart / NormalWishart / Wishart chain (mvnormal.cpp:56-135,
//                       bpmf.h:98-103) as ONE single-block kernel
//   items_exact_kernel  Sys::sample(idx, other) (sample.cpp:263-336) for any K, one CTA per item
//   stats kernels       sums/prods/norms + cov (sample.cpp:359-362,379-384)
//   predict kernels     Sys::predict (sample.cpp:48-96)
#include "common.cuh"
#include <algorithm>
#include "rng.cuh"

namespace bpmf {

constexpr unsigned long long ERR_HYPER_NOT_PD = 5ull << 32;

// =================================================================================================
// Hyper-parameter draw
// =================================================================================================
struct HyperArgs {
    int K, N, nblk;
    int sequential;     // != 0: the plain sequential gamma walk (BPMF_HYPER_SEQUENTIAL: tests of the fall-back path)
    uint32_t iter;
    const double *sum;  // K or nullptr (zeros)
    const double *cov;  // K*K
    uint32_t *words;
    unsigned char *acc;
    int *rank, *pos_of_rank, *row_start, *row_cls, *piv;
    double *mats, *vecs;
    double *mu, *LambdaU, *LambdaF;
    unsigned long long *err;
};

__device__ __forceinline__ double normal_at(const uint32_t *words, int q)
{
    const Polar p = polar_attempt(words[2 * q], words[2 * q + 1], words[2 * q + 2], words[2 * q + 3]);
    return p.y * polar_mult(p.r2);
}

// Where the kernel's work arrays live. The chain is one block's work and every step waits for the previous one, so its
// duration is the LATENCY of its memory round trips: from global scratch (L2) the K = 32 draw takes 0.17 ms, which is the
// critical path of a small problem's sweep (ML-100K: 0.25 ms per sweep). When they fit, the matrices (5 K^2 doubles) and the
// random-stream arrays are placed in shared memory instead; the arithmetic and its order are untouched.
struct HyperLayout {
    size_t mats_off, vecs_off, piv_off, rows_off, rng_off, bytes;   // offsets into dynamic shared memory
    bool mats, rng;
};
__host__ __device__ inline HyperLayout hyper_layout(int K, int nblk, size_t cap)
{
    HyperLayout l{};
    const size_t KK = (size_t)K * K;
    size_t off = 0;
    l.mats = 5 * KK * 8 + 4 * (size_t)K * 8 + 3 * (size_t)(K + 1) * 4 + 64 <= cap;
    if (l.mats) {
        l.mats_off = off; off += 5 * KK * 8;
        l.vecs_off = off; off += 4 * (size_t)K * 8;
        l.piv_off = off; off += (((size_t)K * 4 + 15) / 16) * 16;
        l.rows_off = off; off += ((2 * (size_t)(K + 1) * 4 + 15) / 16) * 16;
    }
    const size_t rng_bytes = (size_t)nblk * 16 + (((size_t)nblk * 2 + 15) / 16) * 16 + 2 * (size_t)nblk * 2 * 4;
    l.rng = l.mats && off + rng_bytes <= cap;
    if (l.rng) { l.rng_off = off; off += rng_bytes; }
    l.bytes = off;
    return l;
}

static int hyper_threads(int K) { return K <= 32 ? 256 : 1024; }   // measured: bench_micro/hyper_timing.py

#ifdef BPMF_HYPER_PROF     // phase probes: clock64 of thread 0 at the phase boundaries, printed by the kernel (bench_micro/hyper_timing.py)
#define HSTAMP(i) do { if (threadIdx.x == 0) hprof[i] = clock64(); } while (0)
#else
#define HSTAMP(i) do { } while (0)
#endif

// SM / SR: the matrices / the random-stream arrays are in shared memory (compile-time, so that the accesses are LDS / STS and
// not generic loads)
template <bool SM, bool SR>
__global__ void __launch_bounds__(1024) hyper_kernel(HyperArgs a, const size_t smem_cap)
{
#ifdef BPMF_HYPER_PROF
    __shared__ long long hprof[16];
#endif
    HSTAMP(0);
    const int tid = threadIdx.x, T = blockDim.x;
    const int K = a.K, N = a.N, nblk = a.nblk;
    const int KK = K * K;
    extern __shared__ __align__(16) unsigned char hyper_smem[];
    const HyperLayout lay = hyper_layout(K, nblk, smem_cap);
    double *Uout = a.LambdaU;                 // where U = LambdaU is kept while the kernel works on it
    if (SM) {
        a.mats = reinterpret_cast<double *>(hyper_smem + lay.mats_off);
        Uout = a.mats + 4 * (size_t)KK;
        a.vecs = reinterpret_cast<double *>(hyper_smem + lay.vecs_off);
        a.piv = reinterpret_cast<int *>(hyper_smem + lay.piv_off);
        a.row_start = reinterpret_cast<int *>(hyper_smem + lay.rows_off);
        a.row_cls = a.row_start + (K + 1);
    }
    if (SR) {
        unsigned char *q = hyper_smem + lay.rng_off;
        a.words = reinterpret_cast<uint32_t *>(q); q += (size_t)nblk * 16;
        a.acc = q; q += (((size_t)nblk * 2 + 15) / 16) * 16;
        a.rank = reinterpret_cast<int *>(q); q += (size_t)nblk * 2 * 4;
        a.pos_of_rank = reinterpret_cast<int *>(q);
    }
    __shared__ int s_scan[1024];
    __shared__ int s_total[2];
    __shared__ int s_piv;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;

    // ---- 1. the raw word stream of rng_set_pos(iter), in MicroURNG delivery order -------------------
    for (int b = tid; b < nblk; b += T) {
        const U4 w = stream_block(a.iter, (uint32_t)b);
        a.words[4 * b + 0] = w.v[3];
        a.words[4 * b + 1] = w.v[2];
        a.words[4 * b + 2] = w.v[1];
        a.words[4 * b + 3] = w.v[0];
    }
    __syncthreads();
    HSTAMP(1);
    // ---- 2. acceptance of a polar attempt starting at every even word offset 2q ---------------------
    // (all consumers take words in pairs, so attempts start at even offsets; q odd = straddles two blocks,
    //  which happens after std::gamma_distribution has drawn its single uniform = half a block)
    const int nq = 2 * nblk - 1;
    for (int q = tid; q < nq; q += T)
        a.acc[q] = polar_attempt(a.words[2 * q], a.words[2 * q + 1], a.words[2 * q + 2], a.words[2 * q + 3]).ok ? 1 : 0;
    __syncthreads();
    HSTAMP(2);
    // ---- 3. per class (q & 1): rank[q] = #accepted attempts before q, pos_of_rank[class][r] = t ------
    for (int c = 0; c < 2; ++c) {
        const int nc = (nq - c + 1) / 2;            // attempts q = 2t + c, t in [0, nc)
        const int chunk = (nc + T - 1) / T;
        const int t0 = min(nc, tid * chunk), t1 = min(nc, t0 + chunk);
        int cnt = 0;
        for (int t = t0; t < t1; ++t) cnt += a.acc[2 * t + c];
        s_scan[tid] = cnt;
        __syncthreads();
        for (int off = 1; off < T; off <<= 1) {     // inclusive Hillis-Steele scan
            const int v = (tid >= off) ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        int r = s_scan[tid] - cnt;                  // exclusive prefix
        if (tid == T - 1) s_total[c] = s_scan[tid];
        for (int t = t0; t < t1; ++t) {
            a.rank[2 * t + c] = r;
            if (a.acc[2 * t + c]) { a.pos_of_rank[c * nblk + r] = t; ++r; }
        }
        __syncthreads();
    }
    HSTAMP(3);
    // ---- 4. sequential walk: K gamma draws with the junk / kept normal runs between them ------------
    // (WishartUnitChol, mvnormal.cpp:64-73; std::gamma_distribution = Marsaglia-Tsang,
    //  /usr/include/c++/13/bits/random.tcc:2353-2394). Only the stream POSITIONS of the kept normals are
    //  recorded here; their values are computed in parallel in step 5.
    double *au = a.mats + 3 * (size_t)KK;
    for (int e = tid; e < KK; e += T) au[e] = 0.0;
    const int nu_c = K + N;  // nu + N, mvnormal.cpp:125
    auto next_attempt = [&](int p) -> int {   // first accepted attempt at/after word offset p -> its q, or -1
        const int q = p >> 1;
        if (q >= nq) return -1;
        const int c = q & 1, r = a.rank[q];
        if (r >= s_total[c]) return -1;
        return 2 * a.pos_of_rank[c * nblk + r] + c;
    };
    auto skip_normals = [&](int p, int n, int *start_rank, int *cls) -> int {  // consume n normals from p
        const int q = p >> 1;
        if (q >= nq) return -1;
        const int c = q & 1, r = a.rank[q];
        *start_rank = r; *cls = c;
        if (n == 0) return p;
        if (r + n - 1 >= s_total[c]) return -1;
        return 2 * (2 * a.pos_of_rank[c * nblk + r + n - 1] + c) + 4;
    };
    // 4a. The walk is sequential only in the stream POSITIONS; the arithmetic of a gamma draw (two logarithms, a square root
    // and a division on one thread: 2 000 cycles each) is not. A draw almost always takes its first normal and its first
    // uniform (Marsaglia-Tsang accepts > 99.9 % at these shape parameters), so thread 0 first walks the positions under that
    // assumption — integer table look-ups only —, then K threads do one draw each; if any draw would have gone round its
    // loop again, the plain sequential walk below redoes everything. Either way the result is the sequential one.
    __shared__ int s_gq[129], s_upos[129], s_spec;
    if (tid == 0) {
        int pos = 0;
        bool bad = 0.5 * (nu_c - (K - 1)) < 1.0;
        for (int i = 0; i < K && !bad; ++i) {
            const int q = next_attempt(pos);
            if (q < 0) { bad = true; break; }
            s_gq[i] = q;
            pos = 2 * q + 4;
            if ((pos >> 1) >= 2 * nblk || pos + 1 >= 4 * nblk) { bad = true; break; }
            s_upos[i] = pos;
            pos += 2;
            int r0, c0;
            pos = skip_normals(pos, K - i - 1, &r0, &c0);
            if (pos < 0) { bad = true; break; }
            pos = skip_normals(pos, K - i - 1, &r0, &c0);
            if (pos < 0) { bad = true; break; }
            a.row_start[i] = r0; a.row_cls[i] = c0;
        }
        if (!bad) {
            int r0, c0;
            pos = skip_normals(pos, K, &r0, &c0);
            if (pos < 0) bad = true;
            a.row_start[K] = r0; a.row_cls[K] = c0;
        }
        s_spec = (bad || a.sequential) ? 0 : 1;      // (bad: let the sequential walk find out what is wrong)
    }
    __syncthreads();
    int spec_ok = s_spec;
    if (spec_ok) {
        bool mine = true;
        if (tid < K) {
            const int i = tid;
            const double alpha = 0.5 * (nu_c - i);
            const double a1 = alpha - 1.0 / 3.0;
            const double a2 = 1.0 / sqrt(9.0 * a1);
            const int q = s_gq[i], up = s_upos[i];
            const Polar p = polar_attempt(a.words[2 * q], a.words[2 * q + 1], a.words[2 * q + 2], a.words[2 * q + 3]);
            const double mult = polar_mult(p.r2);
            const double n = p.y * mult;
            double v = 1.0 + a2 * n;
            if (v <= 0.0) mine = false;
            else {
                v = v * v * v;
                const double u = canonical(a.words[up], a.words[up + 1]);
                const bool again = (u > 1.0 - 0.0331 * n * n * n * n) && (log(u) > (0.5 * n * n + a1 * (1.0 - v + log(v))));
                if (again) mine = false;
                else au[i + i * K] = sqrt(2.0 * (a1 * v * 1.0));
            }
        }
        spec_ok = __syncthreads_and(mine);
    }
    // 4c. the sequential walk (a draw went round its loop again, or the stream ran out)
    if (!spec_ok && tid == 0) {
        int pos = 0;             // word offset, always even
        bool bad = false;
        for (int i = 0; i < K && !bad; ++i) {
            const double alpha = 0.5 * (nu_c - i);
            if (alpha < 1.0) { bad = true; break; }  // would need the pow() branch; N >= 1 never gets here
            const double a1 = alpha - 1.0 / 3.0;
            const double a2 = 1.0 / sqrt(9.0 * a1);
            bool saved_ok = false;
            double saved = 0.0, n, v, u;
            bool again;
            do {
                do {
                    if (saved_ok) { saved_ok = false; n = saved; }
                    else {
                        const int q = next_attempt(pos);
                        if (q < 0) { bad = true; break; }
                        const Polar p = polar_attempt(a.words[2 * q], a.words[2 * q + 1], a.words[2 * q + 2], a.words[2 * q + 3]);
                        const double mult = polar_mult(p.r2);
                        saved = p.x * mult; saved_ok = true;
                        n = p.y * mult;
                        pos = 2 * q + 4;
                    }
                    v = 1.0 + a2 * n;
                } while (v <= 0.0);
                if (bad) break;
                v = v * v * v;
                if ((pos >> 1) >= 2 * nblk - 0 || pos + 1 >= 4 * nblk) { bad = true; break; }
                u = canonical(a.words[pos], a.words[pos + 1]);
                pos += 2;
                again = (u > 1.0 - 0.0331 * n * n * n * n) && (log(u) > (0.5 * n * n + a1 * (1.0 - v + log(v))));
            } while (again);
            if (bad) break;
            au[i + i * K] = sqrt(2.0 * (a1 * v * 1.0));
            int r0, c0;
            pos = skip_normals(pos, K - i - 1, &r0, &c0);   // VectorXd r = nrandn(K-i-1), discarded (:70)
            if (pos < 0) { bad = true; break; }
            pos = skip_normals(pos, K - i - 1, &r0, &c0);   // the kept ones (:71)
            if (pos < 0) { bad = true; break; }
            a.row_start[i] = r0; a.row_cls[i] = c0;
        }
        if (!bad) {
            int r0, c0;
            pos = skip_normals(pos, K, &r0, &c0);           // MvNormalChol_prec: nrandn(K) (mvnormal.cpp:58)
            if (pos < 0) bad = true;
            a.row_start[K] = r0; a.row_cls[K] = c0;
        }
        if (bad) { atomicMax(a.err, ERR_RNG); s_fail = 1; }
    }
    __syncthreads();
    if (s_fail) return;
    HSTAMP(4);
    // ---- 5. values of the kept normals ---------------------------------------------------------------
    double *zv = a.vecs;              // K
    for (int e = tid; e < KK; e += T) {
        const int i = e % K, j = e / K;
        if (j > i) {
            const int c = a.row_cls[i];
            au[i + j * K] = normal_at(a.words, 2 * a.pos_of_rank[c * nblk + a.row_start[i] + (j - i - 1)] + c);
        }
    }
    for (int t = tid; t < K; t += T) {
        const int c = a.row_cls[K];
        zv[t] = normal_at(a.words, 2 * a.pos_of_rank[c * nblk + a.row_start[K] + t] + c);
    }
    HSTAMP(5);
    // ---- 6. CondNormalWishart (mvnormal.cpp:116-125) --------------------------------------------------
    const double kappa = 2.0;
    double *mu_c = a.vecs + K, *mu_m = a.vecs + 2 * K;
    for (int i = tid; i < K; i += T) {
        const double Um = (a.sum ? a.sum[i] : 0.0) / N;
        mu_m[i] = 0.0 - Um;
        mu_c[i] = (kappa * 0.0 + N * Um) / (kappa + N);
    }
    __syncthreads();
    const double kappa_c = kappa + N;
    const double kappa_m = (kappa * N) / (kappa + N);
    double *lu = a.mats;                       // X, then its LU factors
    double *bt = a.mats + (size_t)KK;          // inverse work, layout [i][c]
    double *L = a.mats + 2 * (size_t)KK;       // T_c, then its lower Cholesky factor
    for (int e = tid; e < KK; e += T) {
        const int i = e % K, j = e / K;
        lu[e] = ((i == j ? 1.0 : 0.0) + N * a.cov[e]) + kappa_m * (mu_m[i] * mu_m[j]);
    }
    __syncthreads();
    HSTAMP(6);
    // LU with partial pivoting (first maximum wins, like a sequential strict-> scan)
    for (int k = 0; k < K; ++k) {
        if (tid < 32) {
            double best = -1.0; int bi = K;
            for (int i = k + tid; i < K; i += 32) {
                const double v = fabs(lu[i + k * K]);
                if (v > best) { best = v; bi = i; }
            }
            for (int off = 16; off; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_piv = bi; a.piv[k] = bi; }
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k)
            for (int j = tid; j < K; j += T) {
                const double t0 = lu[k + j * K];
                lu[k + j * K] = lu[p + j * K];
                lu[p + j * K] = t0;
            }
        __syncthreads();
        const double d = lu[k + k * K];
        for (int i = k + 1 + tid; i < K; i += T) lu[i + k * K] = lu[i + k * K] / d;
        __syncthreads();
        const int m = K - k - 1;
        for (int e = tid; e < m * m; e += T) {
            const int i = k + 1 + e % m, j = k + 1 + e / m;
            lu[i + j * K] = lu[i + j * K] - lu[i + k * K] * lu[k + j * K];
        }
        __syncthreads();
    }
    HSTAMP(7);
    // T_c = X^-1: thread c solves column c (not symmetrised, mvnormal.cpp:124)
    for (int c = tid; c < K; c += T) {
        for (int i = 0; i < K; ++i) bt[i * K + c] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < K; ++k) {
            const int p = a.piv[k];
            if (p != k) { const double t0 = bt[k * K + c]; bt[k * K + c] = bt[p * K + c]; bt[p * K + c] = t0; }
        }
        // (the products of a row do not depend on its running sum: unrolled, they and their loads run ahead of the chain
        //  of subtractions, whose order stays j ascending)
        for (int i = 0; i < K; ++i) {
            double s = bt[i * K + c];
#pragma unroll 8
            for (int j = 0; j < i; ++j) s -= lu[i + j * K] * bt[j * K + c];
            bt[i * K + c] = s;
        }
        for (int i = K - 1; i >= 0; --i) {
            double s = bt[i * K + c];
#pragma unroll 8
            for (int j = i + 1; j < K; ++j) s -= lu[i + j * K] * bt[j * K + c];
            bt[i * K + c] = s / lu[i + i * K];
        }
        for (int i = 0; i < K; ++i) L[i + c * K] = bt[i * K + c];
    }
    __syncthreads();
    HSTAMP(8);
    // chol = T_c.llt(): lower factor from the lower triangle (mvnormal.cpp:78)
    for (int k = 0; k < K; ++k) {
        const double x = L[k + k * K];
        if (x <= 0.0) { if (tid == 0) atomicMax(a.err, ERR_HYPER_NOT_PD | (unsigned)k); return; }
        const double sx = sqrt(x);
        for (int i = k + 1 + tid; i < K; i += T) L[i + k * K] = L[i + k * K] / sx;
        __syncthreads();
        if (tid == 0) L[k + k * K] = sx;
        const int m = K - k - 1;
        for (int e = tid; e < m * m; e += T) {
            const int i = k + 1 + e % m, c = k + 1 + e / m;
            if (i >= c) L[i + c * K] = L[i + c * K] - L[i + k * K] * L[c + k * K];
        }
        __syncthreads();
    }
    HSTAMP(9);
    // U = au * chol.matrixU()  (mvnormal.cpp:83)
    double *U = Uout;
    for (int e = tid; e < KK; e += T) {
        const int i = e % K, j = e / K;
        double s = 0.0;
        if (i <= j)
            for (int k = i; k <= j; ++k) s += au[i + k * K] * L[j + k * K];
        U[e] = s;
    }
    __syncthreads();
    HSTAMP(10);
    // mu = U \ z / sqrt(kappa_c) + mu_c  (mvnormal.cpp:58-60), column-oriented back substitution
    for (int j = K - 1; j >= 0; --j) {
        if (tid == 0) zv[j] = zv[j] / U[j + j * K];
        __syncthreads();
        const double bj = zv[j];
        for (int i = tid; i < j; i += T) zv[i] = zv[i] - U[i + j * K] * bj;
        __syncthreads();
    }
    HSTAMP(11);
    const double sk = sqrt(kappa_c);
    for (int i = tid; i < K; i += T) a.mu[i] = (zv[i] / sk) + mu_c[i];
    // LambdaF = U^T U  (bpmf.h:101)
    for (int e = tid; e < KK; e += T) {
        const int i = e % K, j = e / K;
        const int kmax = min(i, j);
        double s = 0.0;
        for (int k = 0; k <= kmax; ++k) s += U[k + i * K] * U[k + j * K];
        a.LambdaF[e] = s;
        if (U != a.LambdaU) a.LambdaU[e] = U[e];
    }
#ifdef BPMF_HYPER_PROF
    __syncthreads();
    if (tid == 0) {
        const long long e = clock64();
        printf("hyper K=%d T=%d cycles: words %lld accept %lld scan %lld gamma-walk %lld normals %lld setup %lld LU %lld inverse %lld chol %lld U %lld backsub %lld mu+LambdaF %lld total %lld\n",
               K, T, hprof[1] - hprof[0], hprof[2] - hprof[1], hprof[3] - hprof[2], hprof[4] - hprof[3], hprof[5] - hprof[4], hprof[6] - hprof[5],
               hprof[7] - hprof[6], hprof[8] - hprof[7], hprof[9] - hprof[8], hprof[10] - hprof[9], hprof[11] - hprof[10], e - hprof[11], e - hprof[0]);
    }
#endif
}

// ahead = false: draw into hp on the context's stream. ahead = true: the draw for the NEXT iteration, on the auxiliary
// stream into hp_next (capi.cu swaps it in when that iteration starts), so it runs under the other side's sweep.
cudaError_t launch_hyper(bpmf_gpu_ctx *c, int side, uint32_t iter, const double *d_sum, const double *d_cov, bool ahead)
{
    SideDev &s = c->side[side];
    const HyperScratch &hs = c->hs[side];
    const HyperDev &out = ahead ? s.hp_next : s.hp;
    HyperArgs a;
    a.K = c->K; a.N = s.num; a.nblk = hs.nblk; a.iter = iter;
    a.sequential = getenv("BPMF_HYPER_SEQUENTIAL") != nullptr;
    a.sum = d_sum; a.cov = d_cov;
    a.words = hs.words; a.acc = hs.acc; a.rank = hs.rank; a.pos_of_rank = hs.pos_of_rank;
    a.row_start = hs.row_start; a.row_cls = hs.row_cls; a.piv = hs.piv;
    a.mats = hs.mats; a.vecs = hs.vecs;
    a.mu = out.mu; a.LambdaU = out.LambdaU; a.LambdaF = out.LambdaF;
    a.err = c->d_err;
    static size_t cap[64];                     // per device: opt-in dynamic shared memory the kernel may use (0 = not asked yet)
    size_t &dcap = cap[c->device & 63];
    if (!dcap) {
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
        dcap = optin > 16384 ? (size_t)optin - 10240 : 1;      // s_scan and the other static arrays of the kernel take ~4.2 KB
        if (getenv("BPMF_HYPER_GLOBAL_SCRATCH")) dcap = 1;       // A/B: everything in global scratch, as before
        if (dcap > 1 && (cudaFuncSetAttribute(hyper_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dcap) != cudaSuccess ||
                         cudaFuncSetAttribute(hyper_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dcap) != cudaSuccess)) {
            cudaGetLastError();
            dcap = 1;
        }
    }
    const HyperLayout lay = hyper_layout(c->K, hs.nblk, dcap);
    // threads: every step of the chain ends in a block-wide barrier, whose cost grows with the warps that take part, and at
    // small K there is little to share out ((K - k - 1)^2 elements per elimination step)
    static int threads_env = -1;
    if (threads_env < 0) { const char *t = getenv("BPMF_HYPER_THREADS"); threads_env = t ? atoi(t) : 0; }
    int threads = threads_env >= 32 && threads_env <= 1024 ? (threads_env / 32) * 32 : hyper_threads(c->K);
    cudaStream_t hstream = ahead ? c->aux_stream : c->stream;
    if (lay.mats && lay.rng) hyper_kernel<true, true><<<1, threads, lay.bytes, hstream>>>(a, dcap);
    else if (lay.mats) hyper_kernel<true, false><<<1, threads, lay.bytes, hstream>>>(a, dcap);
    else hyper_kernel<false, false><<<1, threads, 0, hstream>>>(a, dcap);
    c->launches++;
    return cudaGetLastError();
}

// =================================================================================================
// Per-item conditional update, reference order, any K: one CTA per item
// =================================================================================================
struct ItemArgs {
    int K, from, to;
    uint32_t iter;
    double alpha, mean_rating;
    const int64_t *colptr;
    const int32_t *rowidx;
    const double *val;
    const double *other;   // K x num_other
    double *items;         // K x num
    int npeers;
    double *const *peers;
    const double *mu, *LambdaF;
    const double *propLambda;   // K*K x num per-item prior precisions (-m / -l, sample.cpp:272-277) or nullptr
    unsigned int *work_counter;
    unsigned long long *err;
};

constexpr int EXACT_TILE = 16;  // ratings staged per step

size_t exact_items_smem_bytes(int K) { return sizeof(double) * ((size_t)K * K + 3 * (size_t)K + (size_t)EXACT_TILE * K + EXACT_TILE); }

__global__ void items_exact_kernel(ItemArgs p)
{
    extern __shared__ double sm[];
    const int K = p.K, KK = K * K, tid = threadIdx.x, T = blockDim.x;
    double *MM = sm;
    double *rr = MM + KK;
    double *z = rr + K;
    double *ytile = z + K;
    double *w = ytile + EXACT_TILE * K;
    __shared__ int s_item;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = p.from + (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const int idx = s_item;
        if (idx >= p.to) break;
        // rng_set_pos((idx+1) * num_latent * (iter+1))  (sample.cpp:266), uint32 truncation
        const uint32_t c = (uint32_t)(((long long)idx + 1) * (long long)K * ((long long)p.iter + 1));
        if (tid < 32) warp_randn(c, K, z);
        // hp_LambdaF: the item's propagated posterior precision if there is one (sample.cpp:272-283)
        const double *LF = p.propLambda ? p.propLambda + (size_t)idx * KK : p.LambdaF;
        // rr = hp_LambdaF * hp.mu (sample.cpp:285; the global mu even with a propagated posterior, quirk Q5); MM = 0
        for (int a = tid; a < K; a += T) {
            double s = 0.0;
            for (int j = 0; j < K; ++j) s += LF[a + j * K] * p.mu[j];
            rr[a] = s;
        }
        for (int e = tid; e < KK; e += T) MM[e] = 0.0;
        __syncthreads();
        // computeMuLambda (sample.cpp:251-257), entries of the upper triangle accumulate in rating order
        const int64_t ps = p.colptr[idx], pe = p.colptr[idx + 1];
        for (int64_t t0 = ps; t0 < pe; t0 += EXACT_TILE) {
            const int nr = (int)min((int64_t)EXACT_TILE, pe - t0);
            for (int e = tid; e < nr * K; e += T) {
                const int r = e / K, a = e - r * K;
                ytile[e] = p.other[(size_t)p.rowidx[t0 + r] * K + a];
            }
            for (int r = tid; r < nr; r += T) w[r] = (p.val[t0 + r] - p.mean_rating) * p.alpha;
            __syncthreads();
            for (int e = tid; e < KK; e += T) {
                const int a = e % K, b = e / K;
                if (a <= b) {
                    double acc = MM[e];
                    for (int r = 0; r < nr; ++r) acc += ytile[r * K + a] * ytile[r * K + b];
                    MM[e] = acc;
                }
            }
            for (int a = tid; a < K; a += T) {
                double acc = rr[a];
                for (int r = 0; r < nr; ++r) acc += ytile[r * K + a] * w[r];
                rr[a] = acc;
            }
            __syncthreads();
        }
        // lower triangle of MM = LambdaF + alpha * MM (sample.cpp:297-298); the LLT only reads the lower part
        for (int e = tid; e < KK; e += T) {
            const int a = e % K, b = e / K;
            if (a >= b) MM[a + b * K] = LF[a + b * K] + p.alpha * MM[b + a * K];
        }
        __syncthreads();
        // chol.compute(MM) (sample.cpp:306): right-looking, per-element update order j = 0..k-1 as in the oracle
        bool failed = false;
        for (int k = 0; k < K; ++k) {
            const double x = MM[k + k * K];
            if (x <= 0.0) { failed = true; break; }
            const double sx = sqrt(x);
            for (int i = k + 1 + tid; i < K; i += T) MM[i + k * K] = MM[i + k * K] / sx;
            __syncthreads();
            if (tid == 0) MM[k + k * K] = sx;
            const int m = K - k - 1;
            for (int e = tid; e < m * m; e += T) {
                const int i = k + 1 + e % m, cc = k + 1 + e / m;
                if (i >= cc) MM[i + cc * K] = MM[i + cc * K] - MM[i + k * K] * MM[cc + k * K];
            }
            __syncthreads();
        }
        if (failed) {  // THROWERROR("Cholesky failed") (sample.cpp:308): reported through the error word
            if (tid == 0) atomicMax(p.err, ERR_CHOLESKY | (unsigned)idx);
            continue;
        }
        // chol.matrixL().solveInPlace(rr) (sample.cpp:321)
        for (int j = 0; j < K; ++j) {
            if (tid == 0) rr[j] = rr[j] / MM[j + j * K];
            __syncthreads();
            const double bj = rr[j];
            for (int i = j + 1 + tid; i < K; i += T) rr[i] = rr[i] - MM[i + j * K] * bj;
            __syncthreads();
        }
        // rr += nrandn(K) (sample.cpp:322)
        for (int a = tid; a < K; a += T) rr[a] = rr[a] + z[a];
        __syncthreads();
        // chol.matrixU().solveInPlace(rr) (sample.cpp:323)
        for (int j = K - 1; j >= 0; --j) {
            if (tid == 0) rr[j] = rr[j] / MM[j + j * K];
            __syncthreads();
            const double bj = rr[j];
            for (int i = tid; i < j; i += T) rr[i] = rr[i] - MM[j + i * K] * bj;
            __syncthreads();
        }
        // items().col(idx) = rr (sample.cpp:324) — and into every peer replica (replaces send_item)
        for (int a = tid; a < K; a += T) {
            const double v = rr[a];
            p.items[(size_t)idx * K + a] = v;
            for (int q = 0; q < p.npeers; ++q)
                if (p.peers[q] && p.peers[q] != p.items) p.peers[q][(size_t)idx * K + a] = v;
        }
    }
}

cudaError_t launch_items_exact(bpmf_gpu_ctx *c, int side, uint32_t iter, double alpha)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    ItemArgs p;
    p.K = c->K; p.from = s.from; p.to = s.to; p.iter = iter; p.alpha = alpha; p.mean_rating = s.mean_rating;
    p.colptr = s.colptr; p.rowidx = s.rowidx; p.val = s.val;
    p.other = o.items; p.items = s.items;
    p.npeers = s.npeers; p.peers = s.peers_dev;
    p.mu = s.hp.mu; p.LambdaF = s.hp.LambdaF; p.propLambda = s.propLambda;
    p.work_counter = s.work_counter; p.err = c->d_err;
    cudaError_t e = cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned int), c->stream);
    if (e != cudaSuccess) return e;
    const size_t smem = exact_items_smem_bytes(c->K);
    const int threads = c->K <= 48 ? 128 : 256;
    e = cudaFuncSetAttribute(items_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, items_exact_kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long n = (long long)s.to - s.from;
    long long grid = (long long)item_sms(c, side) * per_sm;
    if (grid > n) grid = n;
    if (grid < 1) return cudaSuccess;
    items_exact_kernel<<<(unsigned)grid, threads, smem, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}

// =================================================================================================
// Sweep reductions: sum, prod (outer products), norm over ALL items, then cov (sample.cpp:359-362,379-384)
// =================================================================================================
template <int NE>
__global__ void __launch_bounds__(256) stats_partial_kernel(const double *__restrict__ items, int N, int K, double *__restrict__ partials, int b0,
                                                            int npeers, double *const *__restrict__ peers)
{
    constexpr int TILE = 32;
    extern __shared__ double sm[];  // TILE * K
    const int tid = threadIdx.x, T = blockDim.x, KK = K * K;
    const int blk_id = b0 + (int)blockIdx.x;                 // block of the fixed STATS_BLOCKS decomposition
    const int chunk = (N + STATS_BLOCKS - 1) / STATS_BLOCKS;
    const int i0 = (int)min((long long)N, (long long)blk_id * chunk), i1 = min(N, i0 + chunk);
    double acc[NE];
    int ea[NE], eb[NE];
#pragma unroll
    for (int n = 0; n < NE; ++n) {
        acc[n] = 0.0;
        const int e = tid + n * T;
        ea[n] = e < KK ? e % K : -1;
        eb[n] = e < KK ? e / K : 0;
    }
    double s1 = 0.0, s2 = 0.0;  // for tid < K: sum of x[tid], sum of x[tid]^2
    for (int t0 = i0; t0 < i1; t0 += TILE) {
        const int nt = min(TILE, i1 - t0);
        __syncthreads();
        for (int e = tid; e < nt * K; e += T) sm[e] = items[(size_t)t0 * K + e];
        __syncthreads();
        for (int t = 0; t < nt; ++t) {
            const double *x = sm + t * K;
#pragma unroll
            for (int n = 0; n < NE; ++n)
                if (ea[n] >= 0) acc[n] += x[ea[n]] * x[eb[n]];
            if (tid < K) { const double v = x[tid]; s1 += v; s2 += v * v; }
        }
    }
    const size_t base = (size_t)blk_id * (KK + K + 1);
    auto put = [&](int e, double v) {                        // own copy + every peer's copy of the partials
        partials[base + e] = v;
        for (int q = 0; q < npeers; ++q) {
            double *dst = peers[q];
            if (dst && dst != partials) dst[base + e] = v;
        }
    };
#pragma unroll
    for (int n = 0; n < NE; ++n)
        if (ea[n] >= 0) put(tid + n * T, acc[n]);
    __syncthreads();
    for (int a = tid; a < K; a += T) { put(KK + a, s1); sm[a] = s2; }  // (K <= 256 = T)
    __syncthreads();
    if (tid == 0) {
        double nn = 0.0;
        for (int a = 0; a < K; ++a) nn += sm[a];
        put(KK + K, nn);
    }
}

// Fixed-order sum of the STATS_BLOCKS partials: eight lanes per element, lane j adds blocks j, j + 8, ... in order, then a
// fixed shuffle tree. The order depends on nothing but STATS_BLOCKS, so the statistics are the same for any GPU count.
__global__ void __launch_bounds__(256) stats_sum_kernel(const double *__restrict__ partials, int K, double *sum, double *prod, double *norm)
{
    const int KK = K * K, W = KK + K + 1;
    const int gt = blockIdx.x * 256 + threadIdx.x, e = gt >> 3, j = gt & 7;
    double s = 0.0;
    if (e < W)
        for (int b = j; b < STATS_BLOCKS; b += 8) s += partials[(size_t)b * W + e];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (e < W && j == 0) {
        if (e < KK) prod[e] = s;
        else if (e < KK + K) sum[e - KK] = s;
        else *norm = s;
    }
}
// cov = (prod - sum sum^T / N) / (N - 1)  (sample.cpp:383-384)
__global__ void __launch_bounds__(256) stats_cov_kernel(int N, int K, const double *__restrict__ sum, const double *__restrict__ prod, double *cov)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e < K * K) {
        const int a = e % K, b = e / K;
        cov[e] = (prod[e] - (sum[a] * sum[b] / N)) / (N - 1);
    }
}

// Cross-GPU barrier over peer memory (replaces the MPI_Barrier / MPI_Win_fence of the MPI back ends and the NCCL all-reduce this
// repository used as one): thread q stores this rank's epoch into rank q's arrival word (NVLink store, after a system-wide
// fence so that everything earlier kernels of this stream wrote — latent columns, statistics blocks — is visible first),
// then waits for rank q's epoch in its own buffer. One launch of one warp; no host involvement, so a whole multi-GPU sweep
// is enqueued by one C call. A peer that does not arrive within ~20 s sets the error word instead of hanging the GPU.
__global__ void peer_barrier_kernel(double *const *peers, int npeers, int me, size_t flag_off, unsigned long long epoch, unsigned long long *err)
{
    const int q = threadIdx.x;
    if (q >= npeers || !peers[q] || q == me) return;
    volatile unsigned long long *remote = reinterpret_cast<unsigned long long *>(peers[q] + flag_off) + me;
    volatile unsigned long long *mine = reinterpret_cast<unsigned long long *>(peers[me] + flag_off) + q;
    __threadfence_system();
    *remote = epoch;
    const long long t0 = clock64();
    while (*mine < epoch) {
        if (clock64() - t0 > 40000000000ll) { atomicMax(err, ERR_BARRIER | (unsigned)q); break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

// kind: which set of arrival words / which epoch counter (BARRIER_LATENTS, BARRIER_STATS): barriers of different kinds run
// on different streams and every rank may interleave them differently. What a barrier orders against the peers' later
// stores into this rank's buffers (the sums having read the statistics blocks, ...) is the caller's business (capi.cu).
cudaError_t launch_peer_barrier(bpmf_gpu_ctx *c, int side, int kind, cudaStream_t stream)
{
    SideDev &s = c->side[side];
    if (s.n_stat_peers < 2 || s.stat_rank < 0) return cudaSuccess;
    const size_t flag_off = (size_t)STATS_BLOCKS * ((size_t)c->K * c->K + c->K + 1) + (size_t)kind * MAX_PEERS;
    ++s.barrier_epoch[kind];
    peer_barrier_kernel<<<1, 32, 0, stream>>>(s.stat_peers_dev, s.n_stat_peers, s.stat_rank, flag_off, s.barrier_epoch[kind], c->d_err);
    c->launches++;
    return cudaGetLastError();
}

int stats32_block_items(int num);   // stream_kernel.cu

int stats_block_items(int K, int num)
{
    if (K == 32) return stats32_block_items(num);
    const int chunk = (num + STATS_BLOCKS - 1) / STATS_BLOCKS;
    return chunk > 0 ? chunk : 1;
}

// Per-block partial sums. With statistics peers set (multi-GPU) only the blocks of this context's item range are
// reduced and each partial is stored into every rank's buffer; the range must then be aligned to stats_block_items.
cudaError_t launch_stats_partial(bpmf_gpu_ctx *c, int side, cudaStream_t stream)
{
    SideDev &s = c->side[side];
    const int K = c->K, KK = K * K;
    int b0 = 0, nb = STATS_BLOCKS;
    if (s.n_stat_peers > 0) {
        const int bi = stats_block_items(K, s.num);
        if (s.from % bi != 0 || (s.to % bi != 0 && s.to != s.num)) return cudaErrorInvalidValue;
        b0 = s.from / bi;
        // the range that ends at the last item also owns the (empty) blocks behind it: every block is written by exactly one rank
        const int b1 = (s.to == s.num) ? STATS_BLOCKS : s.to / bi;
        nb = (s.from < s.to) ? b1 - b0 : 0;
    }
    if (nb < 1) return cudaSuccess;
    if (K == 32) return launch_stats_partial32(c, side, b0, nb, stream);  // tensor-core version, same partial layout (stream_kernel.cu)
    const int ne = (KK + 255) / 256;
    const size_t smem = sizeof(double) * 32 * K;
#define BPMF_STATS_CASE(NE) stats_partial_kernel<NE><<<nb, 256, smem, stream>>>(s.items, s.num, K, s.partials, b0, s.n_stat_peers, s.stat_peers_dev)
    if (ne <= 1) BPMF_STATS_CASE(1);
    else if (ne <= 2) BPMF_STATS_CASE(2);
    else if (ne <= 4) BPMF_STATS_CASE(4);
    else if (ne <= 8) BPMF_STATS_CASE(8);
    else if (ne <= 16) BPMF_STATS_CASE(16);
    else if (ne <= 32) BPMF_STATS_CASE(32);
    else if (ne <= 64) BPMF_STATS_CASE(64);
    else return cudaErrorInvalidValue;
#undef BPMF_STATS_CASE
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_stats_final(bpmf_gpu_ctx *c, int side, cudaStream_t stream)
{
    SideDev &s = c->side[side];
    const int K = c->K, KK = K * K, W = KK + K + 1;
    stats_sum_kernel<<<(W * 8 + 255) / 256, 256, 0, stream>>>(s.partials, K, s.sum, s.prod, s.norm);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    stats_cov_kernel<<<(KK + 255) / 256, 256, 0, stream>>>(s.num, K, s.sum, s.prod, s.cov);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_stats(bpmf_gpu_ctx *c, int side)
{
    const cudaError_t e = launch_stats_partial(c, side, c->stream);
    return e != cudaSuccess ? e : launch_stats_final(c, side, c->stream);
}

// =================================================================================================
// Sys::predict (sample.cpp:48-96): one thread per test entry, Welford update of Pavg/Pm2 (quirk Q4 kept)
// =================================================================================================
__global__ void __launch_bounds__(256) predict_kernel(int K, int64_t nnz, int n, double mean_rating, const int32_t *__restrict__ t_col,
                                                      const int32_t *__restrict__ t_row, const double *__restrict__ t_val,
                                                      const double *__restrict__ items, const double *__restrict__ other, double *pavg,
                                                      double *pm2, double *partials)
{
    double se = 0.0, se_avg = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const double *m = items + (size_t)t_col[p] * K;
        const double *u = other + (size_t)t_row[p] * K;
        double d = 0.0;
        for (int a = 0; a < K; ++a) d += m[a] * u[a];
        const double pred = d + mean_rating;
        const double r = t_val[p];
        se += (r - pred) * (r - pred);
        double avg = pavg[p];
        const double delta = pred - avg;
        avg = (n == 0) ? pred : (avg + delta / n);
        pavg[p] = avg;
        pm2[p] = (n == 0) ? 0.0 : pm2[p] + delta * (pred - avg);
        se_avg += (r - avg) * (r - avg);
    }
    __shared__ double s0[256], s1[256];
    s0[threadIdx.x] = se; s1[threadIdx.x] = se_avg;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if ((int)threadIdx.x < off) { s0[threadIdx.x] += s0[threadIdx.x + off]; s1[threadIdx.x] += s1[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s0[0]; partials[2 * blockIdx.x + 1] = s1[0]; }
}
// fixed-order sum of the per-block partials: thread t adds blocks t, t + 256, ...; then a shared-memory tree
__global__ void __launch_bounds__(256) predict_final_kernel(const double *partials, int nb, double *out)
{
    __shared__ double s0[256], s1[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if ((int)threadIdx.x < off) { s0[threadIdx.x] += s0[threadIdx.x + off]; s1[threadIdx.x] += s1[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s0[0]; out[1] = s1[0]; }
}

cudaError_t launch_predict(bpmf_gpu_ctx *c, int side, int n)
{
    SideDev &s = c->side[side];
    const SideDev &o = c->side[1 - side];
    if (s.nnz_test == 0) return cudaSuccess;
    long long nb = (s.nnz_test + 255) / 256;
    if (nb > s.pred_blocks) nb = s.pred_blocks;
    predict_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(c->K, s.nnz_test, n, s.mean_rating, s.t_col, s.t_rowidx, s.t_val, s.items,
                                                         o.items, s.pavg, s.pm2, s.pred_partials);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    predict_final_kernel<<<1, 256, 0, c->stream>>>(s.pred_partials, (int)nb, s.pred_partials + 2 * (size_t)s.pred_blocks);
    c->launches++;
    return cudaGetLastError();
}

// =================================================================================================
// Posterior aggregation of -o (sample.cpp:364-368): aggrMu.col(i) += r; aggrLambda.col(i) += vec(r r^T)
// One thread per element of the K x K outer product, items of [from, to).
// =================================================================================================
// aggrMu / aggrLambda hold the items of [base, ...) only (the range the context samples).
__global__ void __launch_bounds__(256) aggregate_kernel(int K, int from, int to, int base, const double *__restrict__ items, double *aggrMu,
                                                        double *aggrLambda)
{
    const int KK = K * K;
    const int64_t total = (int64_t)(to - from) * KK;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = from + q / KK;
        const int e = (int)(q % KK), a = e % K, b = e / K;
        const double *r = items + (size_t)i * K;
        const double ra = r[a], rb = r[b];
        aggrLambda[(size_t)(i - base) * KK + e] += ra * rb;
        if (b == 0) aggrMu[(size_t)(i - base) * K + a] += ra;
    }
}

// Sys::finalize_mu_lambda (c++/bpmf.cpp:281-295), one warp per item, in place: cov = (prod - sum sum^T / n) / (n - 1),
// aggrLambda.col(i) = cov^-1, aggrMu.col(i) = sum / n. The inverse is an in-place Gauss-Jordan elimination with partial
// pivoting (rows swapped as PartialPivLU would, columns un-swapped at the end) on the matrix in shared memory; lane j owns
// columns j, j + 32, ... (CPL of them).
template <int CPL>   // columns per lane: ceil(K / 32)
__global__ void finalize_aggregates_kernel(int K, int nitems, int nsamples, double *aggrMu, double *aggrLambda)
{
    extern __shared__ double fsm[];                  // per warp: A (K x K, column-major) | colk (K) | perm (K ints)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, KK = K * K;
    double *A = fsm + (size_t)warp * (KK + 2 * K), *colk = A + KK;
    int *perm = reinterpret_cast<int *>(colk + K);
    for (int64_t i = (int64_t)blockIdx.x * nwarp + warp; i < nitems; i += (int64_t)gridDim.x * nwarp) {
        double *mu = aggrMu + (size_t)i * K, *lam = aggrLambda + (size_t)i * KK;
        for (int e = lane; e < KK; e += 32) {
            const int r = e % K, c = e / K;
            A[e] = (lam[e] - (mu[r] * mu[c] / nsamples)) / (nsamples - 1);
        }
        __syncwarp();
        for (int k = 0; k < K; ++k) {
            // pivot: largest |A(r,k)|, r >= k (the first of equals)
            double best = -1.0;
            int prow = k;
            for (int r = k + lane; r < K; r += 32) {
                const double v = fabs(A[r + (size_t)k * K]);
                if (v > best) { best = v; prow = r; }
            }
            for (int off = 16; off; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int orow = __shfl_xor_sync(0xffffffffu, prow, off);
                if (ob > best || (ob == best && orow < prow)) { best = ob; prow = orow; }
            }
            if (lane == 0) perm[k] = prow;
            if (prow != k) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int j = lane + 32 * q;
                    if (j < K) {
                        double *a = A + (size_t)j * K;
                        const double ta = a[k];
                        a[k] = a[prow]; a[prow] = ta;
                    }
                }
            }
            __syncwarp();
            const double piv = A[k + (size_t)k * K];
            for (int r = lane; r < K; r += 32) colk[r] = A[r + (size_t)k * K];
            __syncwarp();
            for (int r = lane; r < K; r += 32) A[r + (size_t)k * K] = (r == k) ? 1.0 : 0.0;   // column k becomes e_k, then is eliminated like the rest
            __syncwarp();
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int j = lane + 32 * q;
                if (j < K) {
                    double *a = A + (size_t)j * K;
                    const double ak = a[k] / piv;
                    for (int r = 0; r < K; ++r)
                        if (r != k) a[r] -= colk[r] * ak;
                    a[k] = ak;
                }
            }
            __syncwarp();
        }
        for (int k = K - 1; k >= 0; --k) {           // undo the row swaps on the columns of the inverse
            const int pk = perm[k];
            if (pk != k)
                for (int r = lane; r < K; r += 32) {
                    const double tmp = A[r + (size_t)k * K];
                    A[r + (size_t)k * K] = A[r + (size_t)pk * K];
                    A[r + (size_t)pk * K] = tmp;
                }
            __syncwarp();
        }
        for (int e = lane; e < KK; e += 32) lam[e] = A[e];
        for (int r = lane; r < K; r += 32) mu[r] = mu[r] / nsamples;
        __syncwarp();
    }
}

cudaError_t launch_finalize_aggregates(bpmf_gpu_ctx *c, int side, int nsamples)
{
    SideDev &s = c->side[side];
    const int K = c->K, n = s.aggr_to - s.aggr_from;
    if (n <= 0) return cudaSuccess;
    const size_t per_warp = sizeof(double) * ((size_t)K * K + 2 * (size_t)K);
    int nwarp = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
    if (nwarp < 1) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * nwarp;
    int nb = (n + nwarp - 1) / nwarp;
    if (nb > c->sm_count * 8) nb = c->sm_count * 8;
    cudaError_t e;
#define BPMF_FIN_CASE(CPL)                                                                                              \
    e = cudaFuncSetAttribute(finalize_aggregates_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                     \
    finalize_aggregates_kernel<CPL><<<nb, 32 * nwarp, smem, c->stream>>>(K, n, nsamples, s.aggrMu, s.aggrLambda)
    if (K <= 32) { BPMF_FIN_CASE(1); }
    else if (K <= 64) { BPMF_FIN_CASE(2); }
    else if (K <= 96) { BPMF_FIN_CASE(3); }
    else { BPMF_FIN_CASE(4); }
#undef BPMF_FIN_CASE
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_aggregate(bpmf_gpu_ctx *c, int side)
{
    SideDev &s = c->side[side];
    const int64_t total = (int64_t)(s.to - s.from) * c->K * c->K;
    if (total <= 0) return cudaSuccess;
    int64_t nb = (total + 255) / 256;
    const int64_t cap = (int64_t)c->sm_count * 16;
    if (nb > cap) nb = cap;
    if (s.from < s.aggr_from || s.to > s.aggr_to) return cudaErrorInvalidValue;   // the range grew after _enable_aggregation
    aggregate_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(c->K, s.from, s.to, s.aggr_from, s.items, s.aggrMu, s.aggrLambda);
    c->launches++;
    return cudaGetLastError();
}

// read-and-reset of the error word in one atomic
__global__ void fetch_error_kernel(unsigned long long *err) { err[1] = atomicExch(err, 0ull); }
cudaError_t launch_fetch_error(bpmf_gpu_ctx *c)
{
    fetch_error_kernel<<<1, 1, 0, c->stream>>>(c->d_err);
    return cudaGetLastError();
}

// =================================================================================================
// RNG probe
// =================================================================================================
__global__ void debug_randn_kernel(uint32_t c, int n, double *out)
{
    extern __shared__ double sm[];
    warp_randn(c, n, sm);
    __syncwarp();
    for (int i = threadIdx.x; i < n; i += 32) out[i] = sm[i];
}
cudaError_t launch_debug_randn(bpmf_gpu_ctx *c, uint32_t seed, int n, double *d_out)
{
    debug_randn_kernel<<<1, 32, sizeof(double) * n, c->stream>>>(seed, n, d_out);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace bpmf
