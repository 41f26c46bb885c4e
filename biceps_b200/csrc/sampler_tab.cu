// K5, table-driven Metropolis kernel: ONE THREAD PER CHAIN on the critical path, everything else in helper warps.
//
// What a step costs at 4096 chains per GPU is the dependent instruction stream of the warp that owns the chain
// (DESIGN.md section 4).  The speculative kernels (sampler_spec.cu, sampler_ws.cu) spread a chain over L lanes and pay
// 8 shuffles, 11 shared-memory loads and two 5-FMA divisions per step on that stream.  Here the stream is reduced to
//
//     position' = (position + delta) & mask          (packed 5-bit fields, one register for all scalar restraints)
//     t'        = table[restraint][position']        (ONE shared-memory load)
//     E'        = ((x_0 + x_1) + x_2) + ...          (the reference's summation order, x_r = t', the others cached)
//     accept    = E' < E - (ln u + margin)           (fp64 compare against a threshold prepared ahead)
//
// because a nuisance move changes one restraint term and that term, as a function of the restraint's sigma (and gamma)
// index, is TABULATED per chain for the chain's current conformational state: a circular window of 16 entries per scalar
// restraint and 16 x 16 for the gamma restraint, filled by a helper warp that follows the chain's random walk.  The
// entries are the exact terms ((Ndof ln sigma + sse / 2 sigma^2) + cnorm) - ref in the reference's operation order
// (Markstein division == IEEE division, spec_common.cuh), so the chain is bit-identical to every other kernel.
//
// Warps of a CTA (one consumer warp = 32 chains per CTA, one CTA per SM):
//   0  consumer    the critical path above, 8 steps per batch on registers; anything rare only raises a flag and the
//                  batch is redone by a careful step-by-step path: an accept decision inside the guard band of the
//                  threshold test, a position outside its tabulated window, a state-move proposal that a per-state lower
//                  bound of the energy cannot reject (lb[state] = min over all nuisance parameters, built by bcp_create:
//                  on large ensembles a state move is almost never accepted, and then E_new >= lb > E - ln u proves it
//                  without evaluating anything), and an accepted state move (the careful path rebuilds the chain's tables).
//   1,2 generators Philox4x32-10, decode into pre-digested step words {packed deltas, byte selector, table offset},
//                  thresholds ln u + margin, lb + ln u for state moves; one batch each in turn, a ring of NS stages.
//   3  tabulator   follows the positions the consumer publishes per batch, keeps every window around its walker
//                  (extends it ahead of the walk, at most 15 entries per dimension so that a slot is never rewritten
//                  while the consumer may still read it) and publishes the valid rectangles two batches ahead.
//   5  accountant  replays {proposal word, accept bit} into the run-length histograms, acceptance counters, state trace and
//                  snapshots (the batches are cut so that a snapshot is the last step of its batch).
//   (warp 4 exits: it would share the consumer's SM sub-partition.)
// Same RNG stream, same fp64 operation order, same outputs as every other Metropolis kernel (bit for bit).
#include "spec_common.cuh"

namespace bcp {
namespace tabk {

#ifndef TAB_BL
#define TAB_BL 8                   // steps per batch (the accept bits of a batch are one byte)
#endif
#ifndef TAB_NS
#define TAB_NS 4                   // ring stages
#endif
#ifndef TAB_MRG0
#define TAB_MRG0 4                 // half-width of a freshly built window
#endif
#ifndef TAB_MRGMIN
#define TAB_MRGMIN 4               // the tabulator extends a window when the walker is closer than this to its edge ...
#endif
#ifndef TAB_MRG
#define TAB_MRG 6                  // ... to this distance
#endif
// timing experiments only (scripts/tune_tab.sh): results are WRONG with any of these set
#ifndef TAB_T_NOACC
#define TAB_T_NOACC 0              // the accountant copies the batch's words and does nothing else
#endif
#ifndef TAB_T_NOTAB
#define TAB_T_NOTAB 0              // the tabulator never extends a window; the consumer ignores the window check
#endif
#ifndef TAB_T_NOGEN
#define TAB_T_NOGEN 0              // two Philox rounds instead of ten
#endif
#ifndef TAB_STATS
#define TAB_STATS 0                // 1: the consumer counts the reasons of its rollbacks into spec_stats[2..4] (diagnostics)
#endif
#define TAB_MARGIN 7.62939453125e-06        // 2^-17 >> error of logf((float)u) (1 ulp of a value below 37: 2.2e-6)
#define TAB_PAD 0x7fffffffu                 // raw proposal word of a padding position

__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TB_DONE_%=;\n"
        "bra TB_WAIT_%=;\n"
        "TB_DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64u(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64d(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ int wrap1(int v, int n) {           // v in [-n, 2n)
    return v < 0 ? v + n : (v >= n ? v - n : v);
}
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

template <int R, bool HG>
struct TabCfg {
    static constexpr int BL = TAB_BL, NS = TAB_NS;
    static constexpr int RS = HG ? R - 1 : R;               // scalar restraints
    static constexpr int NF = RS + (HG ? 2 : 0);            // position fields: sigma of every restraint, then gamma
    static constexpr int NE = RS * 16 + (HG ? 256 : 0);     // table entries per chain, stored [entry][chain]
    // byte offsets inside a ring stage
    static constexpr int O_W = 0;                           // [BL][32] uint4 {delta scalars, delta gamma restraint, selector | r << 16, table offset}
    static constexpr int O_TX = O_W + BL * 512;             // [BL][32] double2 {ln u + margin | +inf, lb + ln u - 2 margin | +inf}
    static constexpr int O_U = O_TX + BL * 512;             // [BL][32] double  Metropolis uniform (careful path)
    static constexpr int O_DW = O_U + BL * 256;             // [BL][32] raw proposal word (accountant, careful path)
    static constexpr int O_PS = O_DW + BL * 128;            // [32] packed positions after the batch     (consumer -> tabulator)
    static constexpr int O_PN = O_PS + 128;
    static constexpr int O_ACC = O_PN + 128;                // [32] accept bits | rebuilt << 16          (consumer -> helpers)
    static constexpr int O_E = O_ACC + 128;                 // [32] energy after the batch               (consumer -> accountant)
    static constexpr int STAGE = O_E + 256;
    static_assert(BL == 8 && NS >= 2 && RS <= 4, "batch / ring / field shape");
};
// raw proposal word: bit 31 nuisance move {bits 1:0 dsigma + 1, 3:2 dgamma + 1, 6:4 restraint}, else the proposed state;
// packed positions: 7-bit fields in bytes (mod 128; a table slot is position & 15), scalars in bytes 0..RS-1 of the first
// word, the gamma restraint's sigma / gamma in bytes 0 / 1 of the second.

// the batches of a launch: BL positions each, cut so that every snapshot step (and the last step) is the LAST position of
// its batch -- the first batch of a segment is padded at the front; every role iterates the same sequence
struct BatchIter {
    int s, rem, s_end, te;
    bool seg_snap;                                           // the current segment ends on a snapshot step
    __device__ void init(int s_end_, int first_snap, int te_) {
        s = 0; s_end = s_end_; te = te_;
        if (first_snap >= 0) { rem = first_snap + 1; seg_snap = true; }
        else { rem = s_end; seg_snap = false; }
    }
    __device__ bool done() const { return s >= s_end; }
    __device__ int take() const { const int t = rem & (TAB_BL - 1); return t ? t : TAB_BL; }
    // real steps s0 .. s0 + BL - pad - 1 at positions pad .. BL - 1
    __device__ void get(int& s0, int& pad, bool& snap_last) const {
        const int t = take();
        s0 = s; pad = TAB_BL - t; snap_last = seg_snap && t == rem;
    }
    __device__ void advance() {
        const int t = take();
        s += t; rem -= t;
        if (rem == 0) {
            if (seg_snap && te > 0 && s + te - 1 < s_end) rem = te;
            else { rem = s_end - s; seg_snap = false; }
        }
    }
};

template <int R, bool HG>
__global__ void __launch_bounds__(256, 1)
mcmc_tab_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                const __grid_constant__ LaunchArgs A) {
    using F = TabCfg<R, HG>;
    constexpr int BL = F::BL, NS = F::NS, RS = F::RS, NF = F::NF, EB = SPEC_ENT_BYTES;
    constexpr int QG = R - 1;                                // the gamma restraint (if HG)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(8) uint64_t bars[3][NS];            // 0 full, 1 empty, 2 free
    __shared__ __align__(8) uint64_t bars_tab[3];            // tables valid after batch b (b & 1); 2: initial build done

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            mbar_init(smem_u32(&bars[0][k]), 32);            // one generator warp
            mbar_init(smem_u32(&bars[1][k]), 32);            // consumer
            mbar_init(smem_u32(&bars[2][k]), 96);            // tabulator + both accountants have copied the stage's words
        }
        mbar_init(smem_u32(&bars_tab[0]), 32);
        mbar_init(smem_u32(&bars_tab[1]), 32);
        mbar_init(smem_u32(&bars_tab[2]), 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_sigma_table(reinterpret_cast<double*>(smem_raw), T.sig4, T.sig4_bytes, &tma_bar);     // has the block barrier
    if (warp == 4) return;

    const long long n = C.n_chains;
    long long c = (long long)blockIdx.x * 32 + lane;
    const bool live = c < n;                                 // dead lanes shadow the last chain, no stores
    if (!live) c = n - 1;
    const int lam = C.lam[c];
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double* __restrict__ gh4 = HG ? T.gsse_halo4 : nullptr;
    const int G = HG ? T.r[QG].n_gamma : 1;
    const int GS = HG ? T.gh4_stride : 0;
    const int s_end = (int)(A.s_end - A.s_begin);
    const long long burn_l = A.burn - A.s_begin;
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    int first_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        first_snap = ns < s_end ? (int)ns : -1;
    }
    BatchIter it;
    it.init(s_end, first_snap, (int)A.traj_every);

    // ---- shared memory ------------------------------------------------------------------------------
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t tab0 = smem0 + (uint32_t)T.sig4_bytes;                    // [NE][32] doubles
    const uint32_t row0 = tab0 + (uint32_t)F::NE * 256u;                     // [32][GS] current gamma rows (halo 4)
    const uint32_t rec0 = row0 + (HG ? 32u * (uint32_t)GS * 8u : 0u);        // [R][32] {sse, ref} of the current state
    const uint32_t cst0 = rec0 + (uint32_t)R * 512u;                         // [32] energy + logZ of the current state
    const uint32_t ring0 = cst0 + 256u;
    const uint32_t rect0 = ring0 + (uint32_t)(NS * F::STAGE);                // [2][6][32] valid windows, positions (tabulator -> consumer)
    const uint32_t idxp0 = rect0 + 2u * 6u * 128u;                           // [2][R + 1][32] true indices after a batch
    const uint32_t q10 = idxp0 + 2u * (uint32_t)(R + 1) * 128u;              // 64 scalar cells to fill
    const uint32_t q20 = q10 + 256u;                                         // 64 row / column segments to fill
    const uint32_t evq0 = q20 + 512u;                                        // BL x 64 histogram events
    const uint32_t lane8 = 8u * (uint32_t)lane, lane4 = 4u * (uint32_t)lane;
    auto ring_of = [&](int stage) { return ring0 + (uint32_t)stage * (uint32_t)F::STAGE; };
    const uint32_t b_full = smem_u32(&bars[0][0]), b_empty = smem_u32(&bars[1][0]), b_free = smem_u32(&bars[2][0]);
    const uint32_t b_tab = smem_u32(&bars_tab[0]);

    // exact term of restraint k of chain L (its current state) at sigma index i, gamma index gi -- the reference's
    // operation order; restraint 0 carries energy + logZ (the sum starts (C + t_0) + t_1 ...)
    auto cell_value = [&](int L, int k, int i, int gi) {
        const uint32_t e = smem0 + (uint32_t)EB * (uint32_t)(T.ent_off[k] + 2 + i);
        const double nl = lds64(e), dv = lds64(e + 8), y = lds64(e + 16);
        const double2 rv = lds128(rec0 + (uint32_t)k * 512u + 16u * (uint32_t)L);
        double num = rv.x;
        if (HG && k == QG) num = lds64(row0 + ((uint32_t)L * (uint32_t)GS + (uint32_t)(gi + 4)) * 8u);
        double t = __dadd_rn(nl, div_markstein(num, dv, y));
        t = __dadd_rn(t, T.r[k].cnorm);
        t = __dsub_rn(t, rv.y);
        if (k == 0) t = __dadd_rn(t, lds64(cst0 + 8u * (uint32_t)L));
        return t;
    };
    auto cell_addr = [&](int L, int k, int slot) {           // slot: position & 15, or (sigma & 15) | (gamma & 15) << 4
        const int e = (HG && k == QG) ? 16 * RS + slot : 16 * k + slot;
        return tab0 + (uint32_t)e * 256u + 8u * (uint32_t)L;
    };
    // the windows of chain L around positions (ps, pn) whose true indices are idx[] / gi: all lanes of the warp
    auto build_chain = [&](int L, uint32_t ps, uint32_t pn, const int* idx, int gi) {
        constexpr int W0 = 2 * TAB_MRG0 + 1;
        constexpr int NC = RS * W0 + (HG ? W0 * W0 : 0);
        for (int e = lane; e < NC; e += 32) {
            if (e < RS * W0) {
                const int k = e / W0, d = e - k * W0 - TAB_MRG0;
                const int p = (int)((ps >> (8 * k)) & 127u);
                int ik = 0;
#pragma unroll
                for (int q = 0; q < RS; ++q) ik = q == k ? idx[q] : ik;
                sts64d(cell_addr(L, k, (p + d) & 15), cell_value(L, k, wrap1(ik + d, T.r[k].n_sigma), 0));
            } else if (HG) {
                const int e2 = e - RS * W0;
                const int dg = e2 / W0 - TAB_MRG0, dp = e2 - (e2 / W0) * W0 - TAB_MRG0;
                const int p = (int)(pn & 127u), g = (int)((pn >> 8) & 127u);
                sts64d(cell_addr(L, QG, ((p + dp) & 15) | (((g + dg) & 15) << 4)),
                       cell_value(L, QG, wrap1(idx[QG] + dp, T.r[QG].n_sigma), wrap1(gi + dg, G)));
            }
        }
    };
    // the current state's record / gamma row of chain L into shared memory: all lanes of the warp
    auto stage_state = [&](int L, int st, int lamL) {
        if (lane < R) sts128(rec0 + (uint32_t)lane * 512u + 16u * (uint32_t)L,
                             *reinterpret_cast<const double2*>(T.rec + ((size_t)st * R + lane) * 2));
        if (lane == R) sts64d(cst0 + 8u * (uint32_t)L, __dadd_rn(T.energy[(size_t)lamL * T.S + st], T.logZ[lamL]));
        if (HG)
            for (int k = lane; k < GS; k += 32) sts64d(row0 + ((uint32_t)L * (uint32_t)GS + (uint32_t)k) * 8u, gh4[(size_t)st * GS + k]);
    };

    if (warp >= 1 && warp <= 3) {
        // =====================================================================================
        // generators: Philox + decode, batch b by generator b % 3
        // =====================================================================================
        const int g = warp - 1;
        const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
        const uint32_t S = (uint32_t)T.S;
        const uint32_t thresh = T.nuis_thresh;
        const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
        const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
        const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);
        const double* __restrict__ lb = T.lb + (size_t)lam * T.S;
        uint32_t rk0[10], rk1[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            rk0[i] = seed_lo + (uint32_t)i * BCP_PHILOX_W0;
            rk1[i] = seed_hi + (uint32_t)i * BCP_PHILOX_W1;
            asm volatile("" : "+r"(rk0[i]), "+r"(rk1[i]));
        }
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        for (int b = 0, bg = 0; !it.done(); ++b, it.advance(), bg = bg == 2 ? 0 : bg + 1) {
            if (bg != g) continue;
            const int stage = b % NS, fill = b / NS;
            if (fill >= 1) mbar_wait(b_free + 8u * (uint32_t)stage, (uint32_t)(fill - 1) & 1u);
            int s0, pad;
            bool snap_last;
            it.get(s0, pad, snap_last);
            const uint32_t rs = ring_of(stage);
            // all BL steps into registers first (the lower-bound gathers of the state moves are in flight together), then the stores
            uint32_t g_ds[BL], g_dn[BL], g_sel[BL], g_off[BL], g_dw[BL];
            double g_th[BL], g_xs[BL], g_lu[BL], g_u[BL];
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                const bool real = j >= pad;
                const int s = real ? s0 + j - pad : s0;
                const unsigned long long ctr = ctr0 + (unsigned long long)(long long)s;
                uint32_t pc0 = (uint32_t)ctr, pc1 = (uint32_t)(ctr >> 32), pc2 = ch_lo, pc3 = ch_hi;
#pragma unroll
                for (int i = 0; i < (TAB_T_NOGEN ? 2 : 10); ++i) {
                    const uint64_t p0 = (uint64_t)BCP_PHILOX_M0 * pc0;
                    const uint64_t p1 = (uint64_t)BCP_PHILOX_M1 * pc2;
                    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ pc1 ^ rk0[i];
                    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ pc3 ^ rk1[i];
                    pc1 = (uint32_t)p1; pc3 = (uint32_t)p0; pc0 = n0; pc2 = n2;
                }
                const bool nuis = pc0 < thresh;
                const uint64_t pr = (uint64_t)pc1 * (nuis ? (uint32_t)R : S);
                const uint32_t hi = (uint32_t)(pr >> 32);
                const uint64_t z1 = (uint64_t)(uint32_t)pr * 3u;
                const uint64_t z2 = (uint64_t)(uint32_t)z1 * 3u;
                const uint32_t a = (uint32_t)(z1 >> 32);                                     // dsigma + 1
                const uint32_t bb = (HG && hi == (uint32_t)QG) ? (uint32_t)(z2 >> 32) : 1u;  // dgamma + 1
                const double u = spec_u2(pc2, pc3);
                // ln u to ~1e-6: lg2.approx (2 ulp of a value below 32) times ln 2 in fp64; below 2^-31 (5e-10 of the steps) NaN,
                // which no comparison of the fast path passes: the exact rule decides
                const double lud = u < 4.656612873077393e-10 ? __longlong_as_double(0x7ff8000000000000ll)
                                                             : __dmul_rn((double)__log2f((float)u), 0.6931471805599453);
                // branch-free step words; a state move and a padding position leave the chain where it is on the fast path
                const bool nu = real && nuis, sm = real && !nuis;
                const bool scal = !HG || hi < (uint32_t)RS;
                const uint32_t cs = (a - 1u) & 127u, cg = (bb - 1u) & 127u;
                g_ds[j] = (nu && scal) ? cs << (8u * (hi & 3u)) : 0u;
                g_dn[j] = (nu && !scal) ? (cs | (cg << 8)) : 0u;
                g_sel[j] = nu ? ((scal ? (0x7760u | hi) : 0x7754u) | (hi << 16)) : (0x7760u | (7u << 16));
                g_off[j] = nu ? (scal ? hi : (uint32_t)RS) * 4096u : 0u;
                g_dw[j] = !real ? TAB_PAD : (nuis ? (0x80000000u | (hi << 4) | (bb << 2) | a) : hi);
                g_th[j] = nu ? __dadd_rn(lud, TAB_MARGIN) : inf;
                // lb + ln u - 2 margin for a state move (the gather is issued for every lane: L2-resident table)
                g_xs[j] = __ldg(lb + (sm ? hi : 0u));
                g_lu[j] = sm ? __dsub_rn(lud, 2.0 * TAB_MARGIN) : inf;
                g_u[j] = real ? u : 1.0;
            }
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                sts128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane, g_ds[j], g_dn[j], g_sel[j], g_off[j]);
                sts128(rs + F::O_TX + 512u * (uint32_t)j + 16u * (uint32_t)lane, make_double2(g_th[j], __dadd_rn(g_xs[j], g_lu[j])));
                sts64d(rs + F::O_U + 256u * (uint32_t)j + lane8, g_u[j]);
                sts32(rs + F::O_DW + 128u * (uint32_t)j + lane4, g_dw[j]);
            }
            mbar_arrive(b_full + 8u * (uint32_t)stage);
        }
        return;
    }

    if (warp == 7) {
        // =====================================================================================
        // tabulator: keeps every chain's windows around its walkers
        // =====================================================================================
        int nf[NF > 0 ? NF : 1];                             // grid length of every field
        int pfull[NF > 0 ? NF : 1], ifull[NF > 0 ? NF : 1], lo[NF > 0 ? NF : 1], hi[NF > 0 ? NF : 1];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int k = f < R ? f : QG;
            nf[f] = f < R ? T.r[k].n_sigma : G;
            ifull[f] = f < R ? C.isg[(size_t)k * n + c] : C.igm[(size_t)QG * n + c];
            pfull[f] = 0; lo[f] = -TAB_MRG0; hi[f] = TAB_MRG0;
        }
        // ---- prologue: stage the chains' current states, build every window around position 0 ----
        {
            const int st_mine = C.state[c];
            for (int L = 0; L < 32; ++L)
                stage_state(L, __shfl_sync(SPEC_FULL, st_mine, L), __shfl_sync(SPEC_FULL, lam, L));
            __syncwarp();
            for (int L = 0; L < 32; ++L) {
                int idx[R];
#pragma unroll
                for (int k = 0; k < R; ++k) idx[k] = __shfl_sync(SPEC_FULL, ifull[k], L);
                const int gi = HG ? __shfl_sync(SPEC_FULL, ifull[NF - 1], L) : 0;
                build_chain(L, 0u, 0u, idx, gi);
            }
            __syncwarp();
            mbar_arrive(b_tab + 16u);
        }
        const uint32_t lane_lt = (1u << lane) - 1u;
        uint32_t n1 = 0, n2 = 0;
        // fill queued cells: scalars one cell per item, the gamma restraint one row / column segment per item
        auto process1 = [&]() {
            __syncwarp();
            for (uint32_t i = (uint32_t)lane; i < n1; i += 32) {
                const uint32_t w = lds32(q10 + 4u * i);
                const int L = (int)(w & 31u), k = (int)((w >> 5) & 7u), slot = (int)((w >> 8) & 15u), ii = (int)(w >> 12);
                sts64d(cell_addr(L, k, slot), cell_value(L, k, ii, 0));
            }
            __syncwarp();
            n1 = 0;
        };
        auto process2 = [&]() {
            __syncwarp();
            const int ci = lane & 15;
            for (uint32_t i = (uint32_t)(lane >> 4); i < n2; i += 2) {
                const uint2 w = lds64u(q20 + 8u * i);
                const int L = (int)(w.x & 31u), dim = (int)((w.x >> 5) & 1u), fslot = (int)((w.x >> 6) & 15u);
                const int len = (int)((w.x >> 10) & 15u), fidx = (int)(w.x >> 14);
                const int vslot = (int)((w.y + (uint32_t)ci) & 15u), vidx0 = (int)(w.y >> 4);
                if (ci < len) {
                    const int vidx = wrap1(vidx0 + ci, dim ? T.r[QG].n_sigma : G);
                    const int slot = dim ? (vslot | (fslot << 4)) : (fslot | (vslot << 4));
                    sts64d(cell_addr(L, QG, slot), dim ? cell_value(L, QG, vidx, fidx) : cell_value(L, QG, fidx, vidx));
                }
            }
            __syncwarp();
            n2 = 0;
        };
        auto emit1 = [&](bool has, uint32_t item) {
            const uint32_t m = __ballot_sync(SPEC_FULL, has);
            if (has) sts32(q10 + 4u * (n1 + (uint32_t)__popc(m & lane_lt)), item);
            n1 += (uint32_t)__popc(m);
            if (n1 >= 32u) process1();
        };
        auto emit2 = [&](bool has, uint32_t x, uint32_t y) {
            const uint32_t m = __ballot_sync(SPEC_FULL, has);
            if (has) sts64u(q20 + 8u * (n2 + (uint32_t)__popc(m & lane_lt)), x, y);
            n2 += (uint32_t)__popc(m);
            if (n2 >= 32u) process2();
        };

        for (int b = 0; !it.done(); ++b, it.advance()) {
            const int stage = b % NS, fill = b / NS;
            const uint32_t rs = ring_of(stage);
            mbar_wait(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
            const uint32_t ps = lds32(rs + F::O_PS + lane4), pn = lds32(rs + F::O_PN + lane4);
            const bool reb = (lds32(rs + F::O_ACC + lane4) >> 16) & 1u;
            mbar_arrive(b_free + 8u * (uint32_t)stage);
            int lo_n[NF > 0 ? NF : 1], hi_n[NF > 0 ? NF : 1];
            if (TAB_T_NOTAB) { mbar_arrive(b_tab + 8u * (uint32_t)(b & 1)); continue; }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const uint32_t pm = f < RS ? (ps >> (8 * f)) & 127u : (pn >> (8 * (f - RS))) & 127u;
                const int d = (int)((pm - (uint32_t)pfull[f] + 64u) & 127u) - 64;    // |d| <= BL
                pfull[f] += d;
                ifull[f] = wrap1(ifull[f] + d, nf[f]);
                const int p = pfull[f];
                int l = lo[f], h = hi[f];
                if (reb) { l = p - TAB_MRG0; h = p + TAB_MRG0; lo[f] = l; hi[f] = h; }    // rebuilt by the consumer's careful path
                else {
                    // extend towards the walker; old and new window together span at most 16 positions, so a slot that the
                    // consumer may still read (it works with the previous rectangle) is never rewritten
                    if (hi[f] - p < TAB_MRGMIN) h = imax(hi[f], imin(p + TAB_MRG, lo[f] + 15));
                    if (p - lo[f] < TAB_MRGMIN) l = imin(lo[f], imax(p - TAB_MRG, hi[f] - 15));
                    if (h - l > 14) { if (p - l > h - p) l = h - 14; else h = l + 14; }
                }
                lo_n[f] = l; hi_n[f] = h;
            }
            // ---- new cells: scalars ----
#pragma unroll
            for (int f = 0; f < RS; ++f) {
                // positions hi+1 .. hi_n, then lo_n .. lo-1 (either range may be empty)
                const int up0 = imax(hi[f] + 1, lo_n[f]), nup = imax(hi_n[f] - up0 + 1, 0);
                const int dn1 = imin(lo[f] - 1, hi_n[f]), ndn = imax(dn1 - lo_n[f] + 1, 0);
                const int cnt = nup + ndn;
                for (int i = 0; __any_sync(SPEC_FULL, i < cnt); ++i) {
                    const int q = i < nup ? up0 + i : lo_n[f] + (i - nup);
                    const int ii = wrap1(ifull[f] + (q - pfull[f]), nf[f]);
                    emit1(i < cnt, (uint32_t)lane | ((uint32_t)f << 5) | ((uint32_t)(q & 15) << 8) | ((uint32_t)ii << 12));
                }
            }
            if (n1) process1();
            // ---- new cells: the gamma restraint (new sigma columns over the new gamma range, new gamma rows over the new
            //      sigma range; cells in both are filled twice with the same value) ----
            if (HG) {
                constexpr int fs = RS, fg = RS + 1;
#pragma unroll
                for (int dim = 0; dim < 2; ++dim) {
                    const int ff = dim ? fg : fs, fv = dim ? fs : fg;          // fixed / varying field of a segment
                    const int up0 = imax(hi[ff] + 1, lo_n[ff]), nup = imax(hi_n[ff] - up0 + 1, 0);
                    const int dn1 = imin(lo[ff] - 1, hi_n[ff]), ndn = imax(dn1 - lo_n[ff] + 1, 0);
                    const int cnt = nup + ndn;
                    const int len = hi_n[fv] - lo_n[fv] + 1;                   // 1 .. 15
                    const int vidx0 = wrap1(ifull[fv] + (lo_n[fv] - pfull[fv]), nf[fv]);
                    for (int i = 0; __any_sync(SPEC_FULL, i < cnt); ++i) {
                        const int q = i < nup ? up0 + i : lo_n[ff] + (i - nup);
                        const int fi = wrap1(ifull[ff] + (q - pfull[ff]), nf[ff]);
                        emit2(i < cnt, (uint32_t)lane | ((uint32_t)dim << 5) | ((uint32_t)(q & 15) << 6) | ((uint32_t)len << 10) | ((uint32_t)fi << 14),
                              (uint32_t)(lo_n[fv] & 15) | ((uint32_t)vidx0 << 4));
                    }
                }
                if (n2) process2();
            }
            // ---- publish the rectangles (valid two batches on) and the true indices after this batch ----
            uint32_t nls = 0, mgs = 0, nln = 0, mgn = 0;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                lo[f] = lo_n[f]; hi[f] = hi_n[f];
                const uint32_t nl = (uint32_t)(-lo[f]) & 127u, mg = (uint32_t)(127 - (hi[f] - lo[f]));
                if (f < RS) { nls |= nl << (8 * f); mgs |= mg << (8 * f); }
                else { nln |= nl << (8 * (f - RS)); mgn |= mg << (8 * (f - RS)); }
            }
#pragma unroll
            for (int f = RS; f < 4; ++f) mgs |= 127u << (8 * f);              // unused fields: position 0 in window [0, 0]
            if (!HG) mgn = 127u | (127u << 8);
            const uint32_t rb = rect0 + (uint32_t)(b & 1) * 768u + lane4;
            sts32(rb, nls); sts32(rb + 128u, mgs); sts32(rb + 256u, nln); sts32(rb + 384u, mgn);
            sts32(rb + 512u, ps); sts32(rb + 640u, pn);                       // the positions the true indices below belong to
            const uint32_t ib = idxp0 + (uint32_t)(b & 1) * (uint32_t)(R + 1) * 128u + lane4;
#pragma unroll
            for (int k = 0; k < R; ++k) sts32(ib + 128u * (uint32_t)k, (uint32_t)ifull[k]);
            if (HG) sts32(ib + 128u * (uint32_t)R, (uint32_t)ifull[NF - 1]);
            mbar_arrive(b_tab + 8u * (uint32_t)(b & 1));
        }
        return;
    }

    if (warp == 5 || warp == 6) {
        // =====================================================================================
        // accountants: the bookkeeping of the batches the consumer has finished.  Accountant p owns the batches b with
        // b & 1 == p (events, state trace, snapshots) and only follows the indices through the other one's batches
        // =====================================================================================
        const int ap = warp - 5;
        const uint32_t evq = evq0 + (uint32_t)ap * (uint32_t)(BL * 64 * 8);
        const int grp = C.group[c];
        unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
        unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
        const bool tracing = A.state_trace != nullptr;
        const bool traced = tracing && live && c < A.trace_n;
        const uint32_t lane_lt = (1u << lane) - 1u;
        const uint32_t bin0 = (uint32_t)grp * (uint32_t)T.HB;
        int state = C.state[c];
        int isg[R], since_sg[R];
        unsigned int acc_r[R];
        int igm = HG ? C.igm[(size_t)QG * n + c] : 0, since_gm = burn, since_state = burn;
        unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) { isg[k] = C.isg[(size_t)k * n + c]; since_sg[k] = burn; acc_r[k] = 0; }

        for (int b = 0; !it.done(); ++b, it.advance()) {
            const int stage = b % NS, fill = b / NS;
            const uint32_t rs = ring_of(stage);
            const bool mine = (b & 1) == ap;
            int s0, pad;
            bool snap_last;
            it.get(s0, pad, snap_last);
            mbar_wait(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
            uint32_t w_dw[BL];
#pragma unroll
            for (int j = 0; j < BL; ++j) w_dw[j] = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
            const uint32_t accb = lds32(rs + F::O_ACC + lane4);
            const double E_end = lds64(rs + F::O_E + lane8);
            mbar_arrive(b_free + 8u * (uint32_t)stage);
            if (TAB_T_NOACC) continue;
            uint32_t evn = 0;
            // (1) one pass over the steps: per restraint the steps with an accepted non-null sigma move (8 bits each) and their
            //     directions, the same for gamma, the accepted state moves, the acceptance counters
            uint32_t fm = 0, up = 0, fm4 = 0, up4 = 0, fmg = 0, upg = 0, stm = 0, accp = 0, acc4 = 0, acct = 0;
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                const uint32_t dw = w_dw[j];
                const uint32_t acc = (dw != TAB_PAD) ? (accb >> j) & 1u : 0u;
                const uint32_t an = acc & (dw >> 31);
                const uint32_t r = (dw >> 4) & 7u, cs = dw & 3u, cg = (dw >> 2) & 3u;
                acct += acc;
                stm |= (acc & ~(dw >> 31) & 1u) << j;
                const uint32_t es = an & (cs != 1u ? 1u : 0u), us = an & (cs == 2u ? 1u : 0u);
                if (R <= 4) {
                    accp += an << (8u * (r & 3u));
                    fm |= es << (8u * (r & 3u) + (uint32_t)j);
                    up |= us << (8u * (r & 3u) + (uint32_t)j);
                } else {
                    const uint32_t lo4 = r < 4u ? 1u : 0u;
                    accp += (an & lo4) << (8u * (r & 3u));
                    fm |= (es & lo4) << (8u * (r & 3u) + (uint32_t)j);
                    up |= (us & lo4) << (8u * (r & 3u) + (uint32_t)j);
                    acc4 += an & (lo4 ^ 1u);
                    fm4 |= (es & (lo4 ^ 1u)) << j;
                    up4 |= (us & (lo4 ^ 1u)) << j;
                }
                if (HG) {
                    const uint32_t eg = an & (r == (uint32_t)QG ? 1u : 0u);
                    fmg |= (eg & (cg != 1u ? 1u : 0u)) << j;
                    upg |= (eg & (cg == 2u ? 1u : 0u)) << j;
                }
            }
            if (mine) {
                acc_total += acct;
#pragma unroll
                for (int k = 0; k < R; ++k) acc_r[k] += k < 4 ? (accp >> (8 * k)) & 0xffu : acc4;
            }
            // (2) per parameter: its events in step order (most lanes have none or one per batch).  A batch of the other
            //     accountant only moves the index and the start of the current run.
            auto events = [&](uint32_t mask, uint32_t ups, int& idx, int& since, int nn, uint32_t binbase) {
                if (!mine) {
                    if (mask) {
                        const int net = 2 * __popc(mask & ups) - __popc(mask);       // |net| <= 8 < grid length
                        idx = wrap1(idx + net, nn);
                        const int s = s0 + (31 - __clz((int)mask)) - pad;
                        since = s > burn ? s : since;
                    }
                    return;
                }
                while (__any_sync(SPEC_FULL, mask != 0u)) {
                    const int j = __ffs((int)mask) - 1;                     // -1 where the lane has no event left
                    const int s = s0 + j - pad;
                    const bool on = mask != 0u, rec_on = s > burn;
                    const int cnt = (on && rec_on && live) ? s - since : 0;
                    const uint32_t bin = binbase + (uint32_t)idx;
                    if (on) {
                        idx = wrap1(idx + (((ups >> j) & 1u) ? 1 : -1), nn);
                        since = rec_on ? s : since;
                    }
                    const bool ev = cnt > 0;
                    const uint32_t m = __ballot_sync(SPEC_FULL, ev);
                    if (ev) sts64u(evq + 8u * (evn + (uint32_t)__popc(m & lane_lt)), bin, (uint32_t)cnt);
                    evn += (uint32_t)__popc(m);
                    mask &= mask - 1u;
                }
            };
#pragma unroll
            for (int k = 0; k < R; ++k)
                events(k < 4 ? (fm >> (8 * k)) & 0xffu : fm4, k < 4 ? (up >> (8 * k)) & 0xffu : up4, isg[k], since_sg[k], T.r[k].n_sigma,
                       bin0 + (uint32_t)T.r[k].hist_sigma);
            if (HG) events(fmg, upg, igm, since_gm, G, bin0 + (uint32_t)T.r[QG].hist_gamma);
            // (3) accepted state moves (rare on large ensembles) and the optional state trace, in step order
            if ((mine && tracing) || __any_sync(SPEC_FULL, stm != 0u)) {
#pragma unroll
                for (int j = 0; j < BL; ++j) {
                    const uint32_t dw = w_dw[j];
                    const int s = s0 + j - pad;
                    if ((stm >> j) & 1u) {
                        if (mine) ++acc_state;
                        const int i2 = (int)dw;
                        if (i2 != state) {
                            const bool rec_on = s > burn;
                            if (mine && rec_on && live) atomicAdd(g_state_counts + state, (unsigned long long)(s - since_state));
                            if (rec_on) since_state = s;
                            state = i2;
                        }
                    }
                    if (mine && traced && dw != TAB_PAD && s >= burn) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
                }
            }
            if (!mine) { if (snap_last) ++snap_idx; continue; }
            __syncwarp();
            for (uint32_t k = (uint32_t)lane; k < evn; k += 32) {
                const uint2 e = lds64u(evq + 8u * k);
                atomicAdd(C.param_counts + e.x, (unsigned long long)e.y);
            }
            __syncwarp();
            if (snap_last) {
                const size_t o = (size_t)snap_idx * n + c;
                if (live) {
                    A.traj_E[o] = E_end;
                    A.traj_state[o] = state;
                    A.traj_accept[o] = (unsigned char)((accb >> (BL - 1)) & 1u);
#pragma unroll
                    for (int k = 0; k < R; ++k) A.traj_idx[((size_t)snap_idx * T.P + k) * n + c] = isg[k];
                    if (HG) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = igm;
                }
                ++snap_idx;
            }
        }
        // ---- acceptance counters of the own batches; accountant 0 flushes the run-length counters and persists the chain
        //      (E is persisted by the consumer) ----
        if (live) {
            atomicAdd(C.accepted + c, (unsigned long long)acc_total);
            atomicAdd(C.sep + (size_t)R * n + c, (unsigned long long)acc_state);
#pragma unroll
            for (int k = 0; k < R; ++k) atomicAdd(C.sep + (size_t)k * n + c, (unsigned long long)acc_r[k]);
        }
        if (live && ap == 0) {
            if (s_end > since_state) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - since_state));
            C.state[c] = state;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (s_end > since_sg[k]) atomicAdd(g_param_counts + T.r[k].hist_sigma + isg[k], (unsigned long long)(s_end - since_sg[k]));
                C.isg[(size_t)k * n + c] = isg[k];
            }
            if (HG) {
                if (s_end > since_gm) atomicAdd(g_param_counts + T.r[QG].hist_gamma + igm, (unsigned long long)(s_end - since_gm));
                C.igm[(size_t)QG * n + c] = igm;
            }
        }
        return;
    }

    // =========================================================================================
    // consumer: the chains' critical path, one chain per lane
    // =========================================================================================
    double E = C.E[c];
    uint32_t Ps = 0u, Pn = 0u;                                // packed positions (relative to the launch's start)
    double t[R];
    int idx0[R];                                              // true indices at the launch's start (careful path of batch 0)
#pragma unroll
    for (int k = 0; k < R; ++k) idx0[k] = C.isg[(size_t)k * n + c];
    const int gi0 = HG ? C.igm[(size_t)QG * n + c] : 0;
    unsigned int n_batches = 0, n_rollbacks = 0;
    const uint32_t tabl = tab0 + lane8;
    constexpr uint32_t MS = 0x7f7f7f7fu, MN = 0x00007f7fu;
    constexpr uint32_t HS = 0x80808080u, HN = 0x00008080u;
    // windows: valid where ((P + neglo) & mask) + marg has no bit 7 set in any field
    constexpr uint32_t NL0 = (uint32_t)TAB_MRG0 * 0x01010101u, MG0 = (uint32_t)(127 - 2 * TAB_MRG0) * 0x01010101u;
    uint32_t ov_nls = 0, ov_mgs = 0, ov_nln = 0, ov_mgn = 0;  // windows of the chains the careful path has just rebuilt
    bool ov = false;

    // exact terms of proposed state i2 at the chain's indices (restraint 0 carries energy + logZ), reference order
    auto eval_state = [&](int i2, const int* idx, int gi, double* xk) {
        const double C2 = __dadd_rn(energy[i2], logZ);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const double2 rv = *reinterpret_cast<const double2*>(T.rec + ((size_t)i2 * R + k) * 2);
            double num = rv.x;
            if (HG && k == QG) num = gh4[(size_t)i2 * GS + gi + 4];
            const uint32_t e = smem0 + (uint32_t)EB * (uint32_t)(T.ent_off[k] + 2 + idx[k]);
            double tt = __dadd_rn(lds64(e), div_markstein(num, lds64(e + 8), lds64(e + 16)));
            tt = __dadd_rn(tt, T.r[k].cnorm);
            tt = __dsub_rn(tt, rv.y);
            if (k == 0) tt = __dadd_rn(tt, C2);
            xk[k] = tt;
        }
        double En = xk[0];
#pragma unroll
        for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
        return En;
    };

    mbar_wait(b_tab + 16u, 0u);                               // the initial tables
#pragma unroll
    for (int k = 0; k < R; ++k) t[k] = lds64(tabl + (uint32_t)((HG && k == QG) ? 16 * RS : 16 * k) * 256u);

    for (int b = 0; !it.done(); ++b, it.advance()) {
        const int stage = b % NS, fill = b / NS;
        const uint32_t rs = ring_of(stage);
        mbar_wait(b_full + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
        uint32_t nls, mgs, nln, mgn;
        if (b >= 2) {
            mbar_wait(b_tab + 8u * (uint32_t)(b & 1), (uint32_t)((b - 2) >> 1) & 1u);
            const uint32_t rb = rect0 + (uint32_t)(b & 1) * 768u + lane4;
            nls = lds32(rb); mgs = lds32(rb + 128u); nln = lds32(rb + 256u); mgn = lds32(rb + 384u);
        } else {
            nls = NL0; mgs = MG0; nln = NL0 & MN; mgn = MG0 & 0xffffu;
#pragma unroll
            for (int f = RS; f < 4; ++f) { nls &= ~(0xffu << (8 * f)); mgs |= 127u << (8 * f); }
            if (!HG) { nln = 0u; mgn = 127u | (127u << 8); }
        }
        if (ov) { nls = ov_nls; mgs = ov_mgs; nln = ov_nln; mgn = ov_mgn; }
        ov = false;

        // ---- fast batch ----------------------------------------------------------------------------
        const double sv_E = E;
        const uint32_t sv_Ps = Ps, sv_Pn = Pn;
        double sv_t[R];
#pragma unroll
        for (int k = 0; k < R; ++k) sv_t[k] = t[k];
        uint32_t accb = 0, why = 0, lbm = 0;
        bool rare = false;
        double Eh[BL];                                        // energy / positions BEFORE every step (SSA values kept alive: no instructions)
        uint32_t Psh[BL], Pnh[BL];
        uint4 wv[BL];
        double2 tx[BL];
#pragma unroll
        for (int j = 0; j < BL; ++j) {
            wv[j] = lds128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane);
            tx[j] = lds128(rs + F::O_TX + 512u * (uint32_t)j + 16u * (uint32_t)lane);
        }
#pragma unroll
        for (int j = 0; j < BL; ++j) {
            const uint4 w = wv[j];
            const uint32_t Ps2 = (Ps + w.x) & MS, Pn2 = (Pn + w.y) & MN;
            const uint32_t x = __byte_perm(Ps2, Pn2, w.z);
            const uint32_t e = (x & 0xfu) | ((x >> 4) & 0xf0u);
            const double tp = lds64(tabl + w.w + e * 256u);
            const uint32_t bad = ((((Ps2 + nls) & MS) + mgs) & HS) | ((((Pn2 + nln) & MN) + mgn) & HN);
            const int r = (int)(w.z >> 16);
            double xk[R];
#pragma unroll
            for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t[k];
            double En = xk[0];
#pragma unroll
            for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
            const double Ahi = __dsub_rn(E, tx[j].x), Alo = __dadd_rn(Ahi, 2.0 * TAB_MARGIN);
            const bool acc = En < Ahi;
            const bool unsure = !acc && !(En > Alo);                     // guard band, also NaN
            rare = rare || unsure || (!TAB_T_NOTAB && bad != 0u);
            lbm |= !(E < tx[j].y) ? (1u << j) : 0u;                       // a state move the lower bound cannot reject: resolved after the batch
            Eh[j] = E; Psh[j] = Ps; Pnh[j] = Pn;
            if (TAB_STATS) { why |= unsure ? 1u : 0u; why |= bad != 0u ? 2u : 0u; }
            E = acc ? En : E;
#pragma unroll
            for (int k = 0; k < R; ++k) t[k] = acc ? xk[k] : t[k];
            Ps = acc ? Ps2 : Ps;
            Pn = acc ? Pn2 : Pn;
            accb |= acc ? (1u << j) : 0u;
        }
        ++n_batches;
        uint32_t rebw = 0;
        if (!__any_sync(SPEC_FULL, rare) && __any_sync(SPEC_FULL, lbm != 0u)) {
            // ---- state-move proposals the lower bound could not reject: evaluated exactly at the chain's indices of that step
            //      (true indices = the tabulator's last publication + the positions' progress since); almost all are rejected
            //      and the batch stands ----
            int idb[R], gib = gi0;
            uint32_t psb = 0u, pnb = 0u;
#pragma unroll
            for (int k = 0; k < R; ++k) idb[k] = idx0[k];
            if (b >= 2) {
                const uint32_t rb = rect0 + (uint32_t)(b & 1) * 768u + lane4;
                const uint32_t ib = idxp0 + (uint32_t)(b & 1) * (uint32_t)(R + 1) * 128u + lane4;
                psb = lds32(rb + 512u); pnb = lds32(rb + 640u);
#pragma unroll
                for (int k = 0; k < R; ++k) idb[k] = (int)lds32(ib + 128u * (uint32_t)k);
                if (HG) gib = (int)lds32(ib + 128u * (uint32_t)R);
            }
            while (__any_sync(SPEC_FULL, lbm != 0u)) {
                const int j = __ffs((int)lbm) - 1;
                if (lbm) {
                    double Ej = Eh[0];
                    uint32_t psj = Psh[0], pnj = Pnh[0];
#pragma unroll
                    for (int q = 1; q < BL; ++q) { Ej = j == q ? Eh[q] : Ej; psj = j == q ? Psh[q] : psj; pnj = j == q ? Pnh[q] : pnj; }
                    int idj[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        const uint32_t pm = (HG && k == QG) ? (pnj & 127u) - (pnb & 127u) : ((psj >> (8 * k)) & 127u) - ((psb >> (8 * k)) & 127u);
                        idj[k] = wrap1(idb[k] + ((int)((pm + 64u) & 127u) - 64), T.r[k].n_sigma);
                    }
                    const int gj = HG ? wrap1(gib + ((int)(((((pnj >> 8) & 127u) - ((pnb >> 8) & 127u)) + 64u) & 127u) - 64), G) : 0;
                    const uint32_t dw = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
                    const double u = lds64(rs + F::O_U + 256u * (uint32_t)j + lane8);
                    double xk[R];
                    const double En = eval_state((int)dw, idj, gj, xk);
                    if ((En < Ej) || metropolis_accept(__dsub_rn(Ej, En), u)) rare = true;      // accepted after all: the careful path
                    if (TAB_STATS) why |= 4u;
                }
                lbm &= lbm - 1u;
            }
        }
        if (__any_sync(SPEC_FULL, rare)) {
            // ---- careful batch: step by step from the raw proposal words, every rare case handled exactly ----
            ++n_rollbacks;
            if (TAB_STATS) {
                const uint32_t w1 = __ballot_sync(SPEC_FULL, why & 1u), w2 = __ballot_sync(SPEC_FULL, why & 2u), w4 = __ballot_sync(SPEC_FULL, why & 4u);
                if (lane == 0) {
                    if (w1) atomicAdd(C.spec_stats + 2, 1ull);
                    if (w2) atomicAdd(C.spec_stats + 3, 1ull);
                    if (w4) atomicAdd(C.spec_stats + 4, 1ull);
                }
            }
            E = sv_E; Ps = sv_Ps; Pn = sv_Pn;
#pragma unroll
            for (int k = 0; k < R; ++k) t[k] = sv_t[k];
            accb = 0;
            int idx[R], gi = gi0;
#pragma unroll
            for (int k = 0; k < R; ++k) idx[k] = idx0[k];
            if (b >= 1) {
                // the tabulator has finished batch b - 1 (and is now idle until this batch is published): its true indices
                mbar_wait(b_tab + 8u * (uint32_t)((b - 1) & 1), (uint32_t)((b - 1) >> 1) & 1u);
                const uint32_t ib = idxp0 + (uint32_t)((b - 1) & 1) * (uint32_t)(R + 1) * 128u + lane4;
#pragma unroll
                for (int k = 0; k < R; ++k) idx[k] = (int)lds32(ib + 128u * (uint32_t)k);
                if (HG) gi = (int)lds32(ib + 128u * (uint32_t)R);
            }
            bool reb = false;
#pragma unroll 1
            for (int j = 0; j < BL; ++j) {
                const uint32_t dw = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
                const double u = lds64(rs + F::O_U + 256u * (uint32_t)j + lane8);
                const uint4 w = lds128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                const double2 txj = lds128(rs + F::O_TX + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                const bool realj = dw != TAB_PAD;
                const bool nuis = (int)dw < 0;
                const int r = (int)((dw >> 4) & 7u), ds = (int)(dw & 3u) - 1, dg = (int)((dw >> 2) & 3u) - 1;
                double En = E;
                double xk[R];
                int inew = 0, gnew = gi;
                bool acc = false;
                if (realj && nuis) {
                    int ik = 0;
#pragma unroll
                    for (int k = 0; k < R; ++k) ik = r == k ? idx[k] : ik;
                    inew = wrap1(ik + ds, T.r[r].n_sigma);
                    if (HG && r == QG) gnew = wrap1(gi + dg, G);
                    const double tp = cell_value(lane, r, inew, gnew);
#pragma unroll
                    for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t[k];
                    En = xk[0];
#pragma unroll
                    for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
                    acc = (En < E) || metropolis_accept(__dsub_rn(E, En), u);
                }
                // state moves: the lower bound rejects almost all of them; the others are evaluated exactly
                const bool st_eval = realj && !nuis && !(E < txj.y);
                if (__any_sync(SPEC_FULL, st_eval)) {
                    if (st_eval) {
                        En = eval_state((int)dw, idx, gi, xk);
                        acc = (En < E) || metropolis_accept(__dsub_rn(E, En), u);
                    }
                    // an accepted state move: the chain's record / gamma row in shared memory, its tables after the batch
                    uint32_t m = __ballot_sync(SPEC_FULL, st_eval && acc);
                    while (m) {
                        const int L = __ffs((int)m) - 1;
                        m &= m - 1;
                        stage_state(L, __shfl_sync(SPEC_FULL, (int)dw, L), __shfl_sync(SPEC_FULL, lam, L));
                    }
                    if (st_eval && acc) reb = true;
                    __syncwarp();
                }
                if (acc) {
                    E = En;
#pragma unroll
                    for (int k = 0; k < R; ++k) t[k] = xk[k];
                    if (nuis) {
#pragma unroll
                        for (int k = 0; k < R; ++k) idx[k] = r == k ? inew : idx[k];
                        gi = gnew;
                        Ps = (Ps + w.x) & MS;
                        Pn = (Pn + w.y) & MN;
                    }
                    accb |= 1u << j;
                }
            }
            // rebuild the tables of the chains that changed their state, around their positions at the end of the batch
            uint32_t m = __ballot_sync(SPEC_FULL, reb);
            while (m) {
                const int L = __ffs((int)m) - 1;
                m &= m - 1;
                int idl[R];
#pragma unroll
                for (int k = 0; k < R; ++k) idl[k] = __shfl_sync(SPEC_FULL, idx[k], L);
                build_chain(L, __shfl_sync(SPEC_FULL, Ps, L), __shfl_sync(SPEC_FULL, Pn, L), idl, __shfl_sync(SPEC_FULL, gi, L));
            }
            __syncwarp();
            if (reb) {
                rebw = 1u << 16;
                ov = true;
                // window [P - MRG0, P + MRG0] in every field
                ov_nls = ((((MS - Ps) + 0x01010101u) & MS) + (NL0 & MS)) & MS;
                ov_nln = ((((MN - Pn) + 0x00000101u) & MN) + (NL0 & MN)) & MN;
                ov_mgs = MG0; ov_mgn = MG0 & 0xffffu;
#pragma unroll
                for (int f = RS; f < 4; ++f) { ov_nls &= ~(0xffu << (8 * f)); ov_mgs |= 127u << (8 * f); }
                if (!HG) { ov_nln = 0u; ov_mgn = 127u | (127u << 8); }
            }
        }
        // ---- publish the batch, release the stage ----------------------------------------------------
        sts32(rs + F::O_PS + lane4, Ps);
        sts32(rs + F::O_PN + lane4, Pn);
        sts32(rs + F::O_ACC + lane4, accb | rebw);
        sts64d(rs + F::O_E + lane8, E);
        mbar_arrive(b_empty + 8u * (uint32_t)stage);
    }

    if (lane == 0 && n_batches) {
        atomicAdd(C.spec_stats, (unsigned long long)n_batches);
        if (n_rollbacks) atomicAdd(C.spec_stats + 1, (unsigned long long)n_rollbacks);
    }
    if (live) C.E[c] = E;
}

template <int R, bool HG>
static cudaError_t launch_tab_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    using F = TabCfg<R, HG>;
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t smem = (size_t)T.sig4_bytes + (size_t)F::NE * 256 + (HG ? (size_t)32 * T.gh4_stride * 8 : 0) + (size_t)R * 512 + 256 +
                        (size_t)F::NS * F::STAGE + 2 * 6 * 128 + 2 * (size_t)(R + 1) * 128 + 256 + 512 + 2 * (size_t)F::BL * 64 * 8 + 128;
    if (smem + 1024 > (size_t)max_optin) return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(mcmc_tab_kernel<R, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)((C.n_chains + 31) / 32);
    mcmc_tab_kernel<R, HG><<<grid, 256, smem, st>>>(T, C, A);
    return cudaGetLastError();
}

}  // namespace tabk

cudaError_t launch_mcmc_tab(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    if (!T.canonical || !T.div_safe || !T.lb || T.R < 1 || T.S < 1) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
    if (T.R - (hg ? 1 : 0) > 4) return cudaErrorNotSupported;
    // single-step index wraps over windows of up to 15 positions; snapshot spacings that keep the batches full
    for (int q = 0; q < T.R; ++q)
        if (T.r[q].n_sigma < 32 || T.r[q].n_sigma >= (1 << 18) || (T.r[q].has_gamma && (T.r[q].n_gamma < 32 || T.r[q].n_gamma >= (1 << 18))))
            return cudaErrorNotSupported;
    if (A.traj_every > 0 && A.traj_every < 4 * TAB_BL) return cudaErrorNotSupported;
#define BCP_TAB(r) case r: return hg ? tabk::launch_tab_one<r, true>(T, C, A, st) : tabk::launch_tab_one<r, false>(T, C, A, st);
    switch (T.R) {
        BCP_TAB(1) BCP_TAB(2) BCP_TAB(3) BCP_TAB(4)
        case 5: return hg ? tabk::launch_tab_one<5, true>(T, C, A, st) : cudaErrorNotSupported;
        default: return cudaErrorNotSupported;
    }
#undef BCP_TAB
}

}  // namespace bcp
