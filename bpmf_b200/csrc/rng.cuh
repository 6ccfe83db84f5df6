// rng.cuh — device restatement of the reference's random stream (c++/mvnormal.cpp:18-47):
// r123::MicroURNG<r123::Philox4x32> keyed {42,0}, counter {c,0,0,block}, feeding libstdc++'s
// std::normal_distribution (Marsaglia polar) through generate_canonical<double,53>.
#pragma once
#include <cstdint>

namespace bpmf {

struct U4 { uint32_t v[4]; };

// Philox4x32-10 (Random123 philox.h constants). ctr = {c0,c1,c2,c3}, key = {k0,k1}.
__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    U4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// Block `blk` of the stream rng_set_pos(c): MicroURNG ORs the block number into ctr[3]
// (Random123 MicroURNG.hpp) and hands the words out as rdata[3], rdata[2], rdata[1], rdata[0].
// stream_word(blk, i) for i = 0..3 is therefore rdata[3 - i].
__device__ __forceinline__ U4 stream_block(uint32_t c, uint32_t blk) { return philox4x32_10(c, 0u, 0u, blk, 42u, 0u); }

// std::generate_canonical<double,53> on a 32-bit URNG: two calls, first call = low word
// (/usr/include/c++/13/bits/random.tcc:3346-3381).
__device__ __forceinline__ double canonical(uint32_t first, uint32_t second)
{
    const double sum = __dadd_rn((double)first, __dmul_rn((double)second, 4294967296.0));
    double r = __dmul_rn(sum, 5.42101086242752217e-20);  // / 2^64, exact
    if (r >= 1.0) r = 0.99999999999999989;               // nextafter(1, 0)
    return r;
}

// 2 * canonical(first, second) - 1 in three fp64 instructions instead of six, the same bits: second * 2^32 is exact, so one
// FMA rounds first + second * 2^32 exactly like the add; the scalings by 2^-64 and by 2 are exact, so a second FMA rounds
// sum * 2^-63 - 1 exactly like the subtraction; the clamp (the sum rounded up to 2^64 -> nextafter(1, 0)) gives 1 - 2^-52.
__device__ __forceinline__ double canonical_2x_minus_1(uint32_t first, uint32_t second)
{
    const double sum = fma((double)second, 4294967296.0, (double)first);
    const double x = fma(sum, 1.0842021724855044e-19, -1.0);          // 2^-63
    return sum >= 18446744073709551616.0 ? 0.99999999999999978 : x;   // 2^64; 1 - 2^-52
}

struct Polar { double x, y, r2; bool ok; };

// One trip of std::normal_distribution's rejection loop (random.tcc:1826-1833) on four stream words.
__device__ __forceinline__ Polar polar_attempt(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    Polar p;
    p.x = canonical_2x_minus_1(w0, w1);
    p.y = canonical_2x_minus_1(w2, w3);
    p.r2 = __dadd_rn(__dmul_rn(p.x, p.x), __dmul_rn(p.y, p.y));  // no FMA: the reference build has none
    p.ok = !(p.r2 > 1.0 || p.r2 == 0.0);
    return p;
}
// mult = sqrt(-2 log(r2) / r2); the returned normal is y * mult, the saved one x * mult (random.tcc:1835-1838)
__device__ __forceinline__ double polar_mult(double r2) { return sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2)); }

// K normals of the stream rng_set_pos(c), computed by one warp: lane l tries block base+l, accepted
// attempts are numbered by ballot/popcount, so z[n] is exactly the n-th randn() of the reference
// (each randn() builds a fresh std::normal_distribution, so it consumes whole blocks and drops the
// second value, mvnormal.cpp:41-43). z must be writable by all lanes (shared memory); caller syncs.
__device__ __forceinline__ void warp_randn(uint32_t c, int K, double *z)
{
    const unsigned lane = threadIdx.x & 31u;
    int have = 0;
    for (uint32_t base = 0; have < K; base += 32) {
        const U4 b = stream_block(c, base + lane);
        const Polar p = polar_attempt(b.v[3], b.v[2], b.v[1], b.v[0]);
        const unsigned m = __ballot_sync(0xffffffffu, p.ok);
        const int n = have + __popc(m & ((1u << lane) - 1u));
        if (p.ok && n < K) z[n] = __dmul_rn(p.y, polar_mult(p.r2));
        have += __popc(m);
    }
}

}  // namespace bpmf
