// Random123/MicroURNG.hpp — minimal stand-in (see philox.h in this directory). MicroURNG turns a counter-based
// generator into a C++11 uniform random number generator for short streams: the user's counter c0 must leave its last
// word free; block n of the stream is CBRNG(c0 with n in the high bits of the last word, key), and the words of a block
// are handed out LAST TO FIRST. (Semantics as published in Random123's MicroURNG.hpp)
#ifndef BPMF_SHIM_R123_MICROURNG
#define BPMF_SHIM_R123_MICROURNG

#include <cstdint>

namespace r123 {

template <typename CBRNG>
class MicroURNG {
  public:
    typedef CBRNG cbrng_type;
    typedef typename CBRNG::ctr_type ctr_type;
    typedef typename CBRNG::key_type key_type;
    typedef typename CBRNG::ukey_type ukey_type;
    typedef uint32_t result_type;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    MicroURNG(ctr_type c0, ukey_type k0) { reset(c0, k0); }
    void reset(ctr_type c0, ukey_type k0)
    {
        c0_ = c0;
        k_ = k0;
        n_ = 0;
        last_elem_ = 0;
    }
    result_type operator()()
    {
        if (last_elem_ == 0) {
            ctr_type c = c0_;
            c[3] |= n_ << 0;          // 32-bit words, BITS = 32: the block number occupies the whole (free) last word
            rdata_ = b_(c, k_);
            ++n_;
            last_elem_ = 4;
        }
        return rdata_[--last_elem_];
    }

  private:
    cbrng_type b_;
    ctr_type c0_, rdata_;
    key_type k_;
    uint32_t n_;
    int last_elem_;
};

}  // namespace r123

#endif
