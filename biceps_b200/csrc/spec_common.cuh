// Helpers shared by the speculative Metropolis kernels (sampler_spec.cu, sampler_ws.cu): the exact reciprocal-table
// division, pinned shared-memory loads, cp.async wrappers, the generated-step word and the per-chain shared-memory block.
#pragma once
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

// tuning knobs (scripts/tune_spec.sh builds variants; the defaults are the measured best)
#ifndef SPEC_EVAL_FIRST
#define SPEC_EVAL_FIRST 0          // source order inside a step: candidates of step s+1 before / after resolving step s
#endif
#ifndef SPEC_PRED_GAMMA
#define SPEC_PRED_GAMMA 0          // 1: only the gamma lane issues the two gamma-value loads (predicated LDS: the other lanes' dummy
                                   // re-reads of their own record cost bank-conflict wavefronts on the SM's shared-memory pipe)
#endif
#ifndef SPEC_BL
#define SPEC_BL(L) ((L) == 8 ? 8 : 4)     // steps per batch as a function of the lanes per chain
#endif
#ifndef SPEC_MINBLOCKS
#define SPEC_MINBLOCKS 1           // __launch_bounds__ second argument (register cap)
#endif
#ifndef SPEC_DEPTH
#define SPEC_DEPTH 2               // batches between the generation (+ prefetch) of a step and its batch; with 3 the
                                   // cp.async wait moves from the last step of a batch to the batch boundary
#endif
#ifndef SPEC_FINISH_BRANCHFREE
#define SPEC_FINISH_BRANCHFREE 0   // 1: philox_finish without the branch ptxas makes of `nuis ? ... : ...`
#endif
#ifndef SPEC_STEP_FENCE
#define SPEC_STEP_FENCE 0          // 1: compiler-level memory fence between the steps of a batch
#endif
#ifndef SPEC_LDS_VOLATILE
#define SPEC_LDS_VOLATILE 1        // shared-memory loads as volatile asm (pins their order in the PTX)
#endif
#if SPEC_LDS_VOLATILE
#define SPEC_ASM asm volatile
#else
#define SPEC_ASM asm
#endif
#define SPEC_FULL 0xffffffffu
#define SPEC_MARGIN 1.220703125e-4f      // 2^-13: >> fp32 error of ln(u) (< 1e-5) + rounding of (float)d (< 3e-6)
constexpr int SPEC_WIN_BYTES = 80;       // 10 doubles
constexpr int SPEC_ENT_BYTES = 24;       // one sigma entry {Ndof ln sigma, 2 sigma^2, 1/(2 sigma^2)}

// a / b with y = RN(1/b) given (host table).  q0 = RN(a y) is within 1.5 ulp of a/b; one Newton
// correction makes q1 faithful (error < 1 ulp); then r1 = a - b q1 is exactly representable and, by
// Markstein's theorem (y correctly rounded, q1 faithful), RN(q1 + r1 y) == RN(a / b).  No exponent
// range issues for the magnitudes admitted by SamplerTables::div_safe.
__device__ __forceinline__ double div_markstein(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, y, q);
}

__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    SPEC_ASM("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// v = on ? [addr] : v -- a predicated load: lanes with on == false generate no shared-memory traffic
__device__ __forceinline__ double lds64_if(bool on, uint32_t addr, double v) {
    SPEC_ASM("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.shared.f64 %0, [%1];\n}" : "+d"(v) : "r"(addr), "r"((uint32_t)on));
    return v;
}
__device__ __forceinline__ double2 lds128_if(bool on, uint32_t addr, double2 v) {
    SPEC_ASM("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\n@p ld.shared.v2.f64 {%0, %1}, [%2];\n}" : "+d"(v.x), "+d"(v.y) : "r"(addr), "r"((uint32_t)on));
    return v;
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    SPEC_ASM("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Metropolis uniform (2k+1) * 2^-53, k = (w2 << 20) | (w3 >> 12)  (== philox.cuh:u2_from_words)
__device__ __forceinline__ double spec_u2(uint32_t w2, uint32_t w3) {
    const uint32_t hi = 0x3FF00000u | (w2 >> 12);
    const uint32_t lo = (w2 << 20) | (w3 >> 12);
    const double one_k = __hiloint2double((int)hi, (int)lo);
    return __dadd_rn(__dadd_rn(one_k, -1.0), 1.1102230246251565e-16);
}

// One generated step.  dw, the word every lane of the group reads:
//   bit 31       nuisance move (else state move)
//   bits 20+r    restraint r is the one that moves (null moves included: they count as accepted)
//   bits 2q+1:2q dsigma + 1 of restraint q (1 = unchanged), bits 17:16 dgamma + 1 (1 = unchanged)
// istate = proposed state of a state move, lu = ln(u) in fp32, u itself is only needed inside the guard band.
#define SPEC_DW_IDLE 0x00015555u
struct SpecGen {
    uint32_t dw, istate;
    float lu;
    double u;
};

struct SpecLayout {          // byte offsets inside a chain's shared-memory block
    int cur_rec, cur_row, slot0, slot_bytes, slot_c, slot_win, chain_bytes;
};

template <int L>
struct SpecCfg {
    static constexpr int BL = SPEC_BL(L);        // steps per batch (one basic block, one rollback unit)
    static constexpr int NB = BL / L;            // Philox blocks a lane generates per batch
    static constexpr int D = SPEC_DEPTH;         // batches between the generation of a step and its batch
    static constexpr int NBATCH = D + 1;         // batches whose proposal slots are alive (this ... the generated one)
    static constexpr int NSLOT = NBATCH * BL;    // proposal slots per chain
};

__host__ __device__ inline SpecLayout spec_layout(int nslot, int R, bool hg, int gh4_stride) {
    SpecLayout o;
    o.cur_rec = 0;
    o.cur_row = 16 * R;
    o.slot0 = o.cur_row + (hg ? gh4_stride * 8 : 0);
    o.slot_c = 16 * R;
    o.slot_win = 16 * R + 16;
    o.slot_bytes = o.slot_win + (hg ? SPEC_WIN_BYTES : 0);
    int b = o.slot0 + nslot * o.slot_bytes;
    b = (b + 15) & ~15;
    while ((b & 127) != 16) b += 16;      // consecutive chains start 16 B apart modulo the 128-B bank period
    o.chain_bytes = b;
    return o;
}

}  // namespace bcp
