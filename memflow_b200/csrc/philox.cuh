// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), stateless: the counter is the
// GLOBAL event index so results do not depend on the chunking or on the number of GPUs.
#pragma once
#include "common.cuh"

namespace mf {

struct Philox4 { uint32_t x, y, z, w; };

MF_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// uniform in (0,1) from the top 23 random bits: (k + 0.5) / 2^23 = (2k + 1) / 2^24, an odd 24-bit numerator, so the sum
// is exact in fp32 and the variate lies strictly inside (0, 1) for every bit pattern (k = 2^23 - 1 gives 1 - 2^-24).
// (24 bits would need 25 significand bits for k >= 2^23: round-to-even then returns exactly 1.0 for b = 0xFFFFFFFF.)
MF_HD float u01_from_bits(uint32_t b) { return ((float)(b >> 9) + 0.5f) * (1.0f / 8388608.0f); }

}  // namespace mf
