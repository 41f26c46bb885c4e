// Deterministic exp for the Metropolis test.
//
// The accept decision of PosteriorSampler.sample (PosteriorSampler.py:322-326) is
// `u < exp(E_old - E_new)`.  To make chain trajectories reproducible bit for bit between the
// CPU oracle and the GPU, exp is *specified* as the fixed sequence of IEEE fp64 add/mul below
// (no FMA contraction, no library call); oracle/sampler_oracle.py:bcp_exp is the same sequence.
// Domain: -40 <= x <= 0 (below -40 the move can never be accepted because u >= 2^-53).
#pragma once
#include <stdint.h>

namespace bcp {

#define BCP_ACCEPT_CUTOFF (-40.0)

__device__ __forceinline__ double exp_det(double x) {
    const double LOG2E = 0x1.71547652b82fep+0;
    const double LN2_HI = 0x1.62e42fee00000p-1;
    const double LN2_LO = 0x1.a39ef35793c76p-33;
    const double RND = 6755399441055744.0;   // 1.5 * 2^52
    const double t = __dadd_rn(__dmul_rn(x, LOG2E), RND);
    const double k = __dsub_rn(t, RND);
    const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(k, LN2_HI)), __dmul_rn(k, LN2_LO));
    const double c[14] = {0x1.0000000000000p+0,  0x1.0000000000000p+0,  0x1.0000000000000p-1,
                          0x1.5555555555555p-3,  0x1.5555555555555p-5,  0x1.1111111111111p-7,
                          0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16,
                          0x1.71de3a556c734p-19, 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26,
                          0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33};
    double p = c[13];
#pragma unroll
    for (int n = 12; n >= 0; --n) p = __dadd_rn(__dmul_rn(p, r), c[n]);
    const int ki = (int)k;                                   // in [-58, 0]
    const double scale = __longlong_as_double((long long)(ki + 1023) << 52);
    return __dmul_rn(p, scale);
}

// Metropolis test for E_new >= E_old; d = E_old - E_new <= 0 (or NaN -> reject).
// Fast path: an fp32 estimate of exp(d) whose relative error is far below the 1e-3 guard
// band decides all but ~2e-3 of the cases; inside the band the exact rule is evaluated, so
// the decision is always the one `u < exp_det(d)` gives.
__device__ __forceinline__ bool metropolis_accept(double d, double u) {
    if (!(d >= BCP_ACCEPT_CUTOFF)) return false;             // also rejects NaN
    const float ef = __expf((float)d);
    const float uf_lo = (float)u * 0.999f;
    const float uf_hi = (float)u * 1.001f;
    if (uf_hi < ef && ef > 1e-30f) return true;
    if (uf_lo > ef) return false;
    return u < exp_det(d);
}

}  // namespace bcp
