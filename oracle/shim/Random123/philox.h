// Random123/philox.h — minimal stand-in for the one generator the reference uses (r123::Philox4x32, 10 rounds), so that
// c++/mvnormal.cpp compiles unmodified with -DBPMF_RANDOM123 on an image without Random123. The round function is the
// published Philox4x32 (Salmon et al., SC'11); tests/test_oracle_kat.py checks it against the Random123 kat_vectors.
// Test infrastructure only.
#ifndef BPMF_SHIM_R123_PHILOX
#define BPMF_SHIM_R123_PHILOX

#include <cstdint>

struct r123array4x32 {
    uint32_t v[4];
    uint32_t &operator[](int i) { return v[i]; }
    const uint32_t &operator[](int i) const { return v[i]; }
};
struct r123array2x32 {
    uint32_t v[2];
    uint32_t &operator[](int i) { return v[i]; }
    const uint32_t &operator[](int i) const { return v[i]; }
};

namespace r123 {

struct Philox4x32 {
    typedef r123array4x32 ctr_type;
    typedef r123array2x32 key_type;
    typedef r123array2x32 ukey_type;
    ctr_type operator()(ctr_type c, key_type k) const
    {
        for (int round = 0; round < 10; ++round) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
            const ctr_type n = {{(uint32_t)(p1 >> 32) ^ c[1] ^ k[0], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], (uint32_t)p0}};
            c = n;
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        return c;
    }
};

}  // namespace r123

#endif
