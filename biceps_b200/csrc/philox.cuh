// Philox4x32-10 and the per-step draw budget of the sampler.
// Executable specification: oracle/philox.py (must agree bit for bit).
#pragma once
#include <stdint.h>

namespace bcp {

#define BCP_PHILOX_M0 0xD2511F53u
#define BCP_PHILOX_M1 0xCD9E8D57u
#define BCP_PHILOX_W0 0x9E3779B9u
#define BCP_PHILOX_W1 0xBB67AE85u

struct Philox4 {
    uint32_t w0, w1, w2, w3;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)BCP_PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)BCP_PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += BCP_PHILOX_W0;
        k1 += BCP_PHILOX_W1;
    }
    Philox4 r;
    r.w0 = c0; r.w1 = c1; r.w2 = c2; r.w3 = c3;
    return r;
}

// block of (chain, step): counter = (step_lo, step_hi, chain_lo, chain_hi), key = seed
__host__ __device__ __forceinline__ Philox4 step_block(uint64_t seed, uint64_t chain, uint64_t step) {
    return philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)chain,
                         (uint32_t)(chain >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Metropolis uniform on the odd 2^-53 grid: never 0, never 1, always an exact double.
__host__ __device__ __forceinline__ double u2_from_words(uint32_t w2, uint32_t w3) {
    const uint64_t k = ((uint64_t)w2 << 20) | (uint64_t)(w3 >> 12);
    return (double)(int64_t)(2 * k + 1) * 1.1102230246251565e-16;   // * 2^-53 (exact)
}

#define BCP_INIT_STEP 0xFFFFFFFFFFFFFFFFull

}  // namespace bcp
