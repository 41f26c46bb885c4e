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
// Warps of a CTA (448 threads; one consumer warp = 32 chains per CTA, one CTA per SM).  A single warp issues dependent integer /
// fp64 code at 0.2-0.4 instructions per cycle, so every role is sized by the length of its dependent chains, the loops of the
// helpers are ROLLED (six code paths share an SM's instruction caches, 128 SMs refill them from L2), and the work that does not
// have to be serial is spread over more warps:
//   0            consumer    the critical path above, 8 steps per batch, registers only; anything rare only raises a flag: a
//                            state-move proposal that a per-state lower bound of the energy cannot reject (lb[state] = min over
//                            all nuisance parameters, built by bcp_create: on large ensembles a state move is almost never
//                            accepted, and then E_new >= lb > E - ln u proves it without evaluating anything) is evaluated
//                            exactly after the batch, and the batch stands unless the move is accepted to another state; an
//                            accept decision inside the guard band of the threshold test, a position outside its tabulated
//                            window, an accepted state move redo the batch with the same step, a vote after each test and the
//                            exact evaluation in place for the lanes concerned (an accepted state move rebuilds the chain's tables).
//   1, 2, 3, 11  generators  Philox4x32-10, decode into pre-digested step words {packed deltas, byte selector, table offset},
//                            thresholds ln u + margin, lb + ln u for state moves, the batch's proposals as bit masks for the
//                            accountants; every batch by all four, two steps each (the ring has to cover a batch's latency).
//   10           tabulator   follows the positions the consumer publishes per batch, keeps every window around its walker
//                            (extends it ahead of the walk; old and new windows of three consecutive batches span at most the
//                            slots of a table, so a slot is never rewritten while the consumer may still read it), publishes the
//                            windows and a checkpoint {positions, true indices}, triple-buffered.
//   5, 6, 7, 9, 13  accountants, one parameter each: accepted non-null moves = proposal masks AND accept bits; run-length
//                            histogram events into a lane-private 16-bin window in shared memory (plain read-modify-write; one
//                            RED per event from every chain of a lambda to the same hot bins is a global rate limit), acceptance
//                            counters, state trace and snapshots (the batches are cut so that a snapshot is the last step of
//                            its batch).
//   4, 8, 12     exit at once: they would share the consumer's SM sub-partition.
// Same RNG stream, same fp64 operation order, same outputs as every other Metropolis kernel (bit for bit).
#include <type_traits>
#include "spec_common.cuh"

namespace bcp {
namespace tabk {

#ifndef TAB_BL
#define TAB_BL 8                   // steps per batch (the accept bits of a batch are one byte)
#endif
#ifndef TAB_NS
#define TAB_NS 4                   // ring stages
#endif
#ifndef TAB_MRG0S
#define TAB_MRG0S 15               // half-width of a freshly built window of a scalar restraint (32 slots: up to 15)
#endif
#ifndef TAB_MRG0
#define TAB_MRG0 6                 // half-width of a freshly built window (gamma restraint: both dimensions)
#endif
#ifndef TAB_MRGMIN
#define TAB_MRGMIN 3               // the tabulator extends a window when the walker is closer than this to its edge ...  (measured 2 / 3 /
                                   // 4 / 5: C3 13.31 / 13.39 / 13.58 / 13.9 ms, C2 1.43 / 1.57 / 1.57 / 1.55e10 -- fewer window moves
                                   // are worth more than the extra proposals outside a window: 25 / 16 / 13 thousand careful batches
                                   // in 4.8 million at 3 / 4 / 5)
#endif
#ifndef TAB_MRG_S
#define TAB_MRG_S 9                // ... to this distance (scalar restraints: 32 slots, windows of up to 31 positions)
#endif
#ifndef TAB_MRG_N
#define TAB_MRG_N 6                // ... (gamma restraint: 16 x 16 slots, windows of up to 15 x 15 positions)
#endif
// timing experiments only (scripts/tune_tab.sh): results are WRONG with any of these set
#ifndef TAB_T_NOACC
#define TAB_T_NOACC 0              // the accountant copies the batch's words and does nothing else
#endif
#ifndef TAB_T_NOTAB
#define TAB_T_NOTAB 0              // the tabulator never extends a window; the consumer ignores the window check
#endif
#ifndef TAB_SPIN
#define TAB_SPIN 0                 // 1: the consumer polls its barriers (test_wait) instead of try_wait; 2: every warp does
#endif
#ifndef TAB_DBG
#define TAB_DBG 0                  // debugging: 1 no self-proposal shortcut, 2 the careful path evaluates every term exactly, 4 no lower-bound resolution
#endif
#ifndef TAB_INPLACE
#define TAB_INPLACE 0              // 1: a proposal outside its tabulated window is evaluated in place (a vote per step); 0: the batch rolls back
#endif
#ifndef TAB_TRACE
#define TAB_TRACE 0                // 1: CTA 0 records clock values of its warps' hand-offs for batches 1000 .. 1039 into spec_stats[8 ...]
#endif
#ifndef TAB_TRACE_B0
#define TAB_TRACE_B0 1000
#endif
#ifndef TAB_T_AFREE
#define TAB_T_AFREE 0              // the accountants do not hold the ring's stages (they may read overwritten words)
#endif
#ifndef TAB_T_NOGEN
#define TAB_T_NOGEN 0              // two Philox rounds instead of ten
#endif
#ifndef TAB_GU
#define TAB_GU 2                   // unroll factor of the generators' step loop (2: two independent Philox chains in flight)
#endif
#ifndef TAB_NG
#define TAB_NG 3                   // generator warps (3 or 4: warp 11 joins)
#endif
#ifndef TAB_CELLPAR
#define TAB_CELLPAR 1    // the tabulator's new scalar cells: one (chain, restraint) at a time across the lanes (0: a queue of cells)
#endif
#ifndef TAB_REPLAY
#define TAB_REPLAY 0     // 1: an accepted state move found after a fast batch: that chain's remaining steps replayed by its lane alone
                         // instead of a careful batch.  Built, bit-identical (90 fuzz cases), 6 x fewer careful batches on C3 (8 300
                         // cycles per replay against 14 500 per careful batch) -- and no faster: C3 13.85 against 13.57 ms per 100 000
                         // steps, C2 1.64e10 against 1.58e10, end-to-end kernels 16.93 against 16.89 ms.  What the rare paths cost is
                         // less their own cycles than the instruction-cache lines they evict for the 14 warps' loops
#endif
#ifndef TAB_NI_U
#define TAB_NI_U 0       // 1: the rare paths' Philox redraw out of line (C3 +-1 %, C2 -6 %: the careful path pays the calls)
#endif
#ifndef TAB_NI_A
#define TAB_NI_A 0       // 1: the rare paths' exact acceptance rule out of line (as above)
#endif
#ifndef TAB_SLIDE
#define TAB_SLIDE 1       // accountants: a sliding 16-bin histogram window (one bin emptied per move at its edge) instead of a re-centred one (all 16 flushed)
#endif
#ifndef TAB_SPEC2
#define TAB_SPEC2 0                // 1: the consumer looks the next step's proposal up under both outcomes of the current step (the address
                                   // computation and the load leave the dependent chain).  Bit-identical, measured: 14.28 ms against 14.16 ms per
                                   // 4096 x 100 000 steps -- 118 instead of 95 instructions per step, and the loop is bound by what one warp
                                   // issues (~2.4 cycles per instruction), not by the chain
#endif
#ifndef TAB_CU
#define TAB_CU 1                   // unroll factor of the consumer's step loop (1: the body stays in the 6 KB L0 instruction cache)
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
// the same as a polling loop (mbarrier.test_wait does not suspend the warp)
__device__ __forceinline__ void mbar_spin(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TB_SPIN_%=:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TB_SPUN_%=;\n"
        "bra TB_SPIN_%=;\n"
        "TB_SPUN_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
// one non-blocking test of a phase (acquire when it has completed)
__device__ __forceinline__ bool mbar_test(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(a), "r"(parity)
        : "memory");
    return ok != 0u;
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

template <int R, bool HG, int NSV = TAB_NS>
struct TabCfg {
    static constexpr int BL = TAB_BL, NS = NSV;
    static constexpr int RS = HG ? R - 1 : R;               // scalar restraints
    static constexpr int NF = RS + (HG ? 2 : 0);            // position fields: sigma of every restraint, then gamma
    static constexpr int NE = RS * 32 + (HG ? 256 : 0);     // table entries per chain, stored [entry][chain]
    // byte offsets inside a ring stage
    static constexpr int O_W = 0;                           // [BL][32] uint4 {delta scalars, delta gamma restraint, selector | r << 16, table offset}
    static constexpr int O_TX = O_W + BL * 512;             // [BL][32] double2 {ln u + margin | +inf, lb + ln u - 2 margin | +inf}
    static constexpr int O_DW = O_TX + BL * 512;            // [BL][32] raw proposal word (accountant, careful path)
    static constexpr int O_MK = O_DW + BL * 128;            // [4]{[32] uint4 + [32] u32}: the batch's proposal masks, one part per generator
    static constexpr int O_PS = O_MK + 4 * (512 + 128);     // [32] packed positions after the batch     (consumer -> tabulator)
    static constexpr int O_PN = O_PS + 128;
    static constexpr int O_ACC = O_PN + 128;                // [32] accept bits | rebuilt << 16          (consumer -> helpers)
    static constexpr int O_E = O_ACC + 128;                 // [32] energy after the batch               (consumer -> accountant)
    static constexpr int STAGE = O_E + 256;
    static_assert(BL == 8 && NS >= 2 && RS <= 4 && TAB_MRGMIN <= TAB_MRG0 && TAB_MRGMIN <= TAB_MRG0S && TAB_MRG0S <= 15, "batch / ring / field shape");
};
// raw proposal word: bit 31 nuisance move {bits 1:0 dsigma + 1, 3:2 dgamma + 1, 6:4 restraint}, else the proposed state;
// packed positions: 7-bit fields in bytes (mod 128; a table slot is position & 31 for a scalar restraint, & 15 in either
// dimension of the gamma restraint), scalars in bytes 0..RS-1 of the first
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

// the consumer's rare paths draw a step's uniform again and apply the exact acceptance rule: out of line, one copy each (the
// rare paths' code is fetched through the same instruction caches as the 14 warps' loops)
__device__ __noinline__ double step_uniform(unsigned long long ctr, unsigned long long chain_id, unsigned long long seed) {
    const Philox4 w = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)chain_id, (uint32_t)(chain_id >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    return spec_u2(w.w2, w.w3);
}
__device__ __noinline__ bool accept_exact(double d, double u) { return bcp::metropolis_accept(d, u); }

// an accountant's 16-bin histogram window of one chain into the global histogram (a walker has left the window): rare, and
// kept out of line -- the accountants' loop has to stay small for the instruction cache
__device__ __noinline__ void flush_window(uint32_t hw, int wb, int nn, unsigned long long* bins) {
#pragma unroll 1
    for (int b4 = 0; b4 < 16; b4 += 4) {                                     // (four loads in flight: the walk is latency, not work)
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = lds32(hw + (uint32_t)(b4 + u) * 128u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u]) {
                int gi = wb + b4 + u;
                gi -= gi >= nn ? nn : 0;
                atomicAdd(bins + gi, (unsigned long long)v[u]);
                sts32(hw + (uint32_t)(b4 + u) * 128u, 0u);
            }
        }
    }
}

template <int R, bool HG, int NSV>
__global__ void __launch_bounds__(448, 1)
mcmc_tab_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                const __grid_constant__ LaunchArgs A) {
    using F = TabCfg<R, HG, NSV>;
    constexpr int BL = F::BL, NS = F::NS, RS = F::RS, NF = F::NF, EB = SPEC_ENT_BYTES;
    constexpr int QG = R - 1;                                // the gamma restraint (if HG)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(8) uint64_t bars[3][NS];            // 0 full, 1 empty, 2 free
    __shared__ __align__(8) uint64_t bars_tab[4];            // tables valid after batch b (b % 3); 3: initial build done

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            mbar_init(smem_u32(&bars[0][k]), 128);           // four generator warps
            mbar_init(smem_u32(&bars[1][k]), 32);            // consumer
            mbar_init(smem_u32(&bars[2][k]), TAB_T_AFREE ? 32 : (HG ? 192 : 160));           // tabulator + four / five accountants have copied the stage's words
        }
        mbar_init(smem_u32(&bars_tab[0]), 32);
        mbar_init(smem_u32(&bars_tab[1]), 32);
        mbar_init(smem_u32(&bars_tab[2]), 32);
        mbar_init(smem_u32(&bars_tab[3]), 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_sigma_table(reinterpret_cast<double*>(smem_raw), T.sig4, T.sig4_bytes, &tma_bar);     // has the block barrier
    if (warp == 4 || warp == 8 || warp == 12 || (warp == 13 && !HG)) return;     // (4, 8, 12 would share the consumer's SM sub-partition)

    const long long n = C.n_chains;
    long long c = (long long)blockIdx.x * 32 + lane;
    const bool live = c < n;                                 // dead lanes shadow the last chain, no stores
    // (diagnostics) event e of warp `warp` at batch b: spec_stats[8 + ((b - TAB_TRACE_B0) * 14 + warp) * 4 + e]
    auto trace = [&](int b, int e) {
        if (TAB_TRACE && blockIdx.x == 0 && lane == 0 && b >= TAB_TRACE_B0 && b < TAB_TRACE_B0 + 40)
            C.spec_stats[8 + ((b - TAB_TRACE_B0) * 14 + warp) * 4 + e] = (unsigned long long)clock64();
    };
    auto trace_t = [&](int b, int k) {                      // (the tabulator's finer points go into the slots of the idle warps 4, 8, 12)
        if (TAB_TRACE && blockIdx.x == 0 && lane == 0 && b >= TAB_TRACE_B0 && b < TAB_TRACE_B0 + 40)
            C.spec_stats[8 + ((b - TAB_TRACE_B0) * 14 + 4 + 4 * (k >> 2)) * 4 + (k & 3)] = (unsigned long long)clock64();
    };
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
    const uint32_t rect0 = ring0 + (uint32_t)(NS * F::STAGE);                // [3][6][32] valid windows, positions (tabulator -> consumer)
    const uint32_t idxp0 = rect0 + 3u * 6u * 128u;                           // [3][R + 1][32] true indices at the checkpoint
    const uint32_t q10 = idxp0 + 3u * (uint32_t)(R + 1) * 128u;              // 64 scalar cells to fill
    const uint32_t q20 = q10 + 256u;                                         // 64 row / column segments to fill
    const uint32_t acc0 = q20 + 512u;                                        // accountants: [R + 1][16][32] histogram windows, 32 dummy words
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
    auto cell_addr = [&](int L, int k, int slot) {           // slot: position & 31, or (sigma & 15) | (gamma & 15) << 4
        const int e = (HG && k == QG) ? 32 * RS + slot : 32 * k + slot;
        return tab0 + (uint32_t)e * 256u + 8u * (uint32_t)L;
    };
    // the windows of chain L around positions (ps, pn) whose true indices are idx[] / gi: all lanes of the warp
    auto build_chain = [&](int L, uint32_t ps, uint32_t pn, const int* idx, int gi) {
        constexpr int W0 = 2 * TAB_MRG0 + 1, W0S = 2 * TAB_MRG0S + 1;
        constexpr int NC = RS * W0S + (HG ? W0 * W0 : 0);
        for (int e = lane; e < NC; e += 32) {
            if (e < RS * W0S) {
                const int k = e / W0S, d = e - k * W0S - TAB_MRG0S;
                const int p = (int)((ps >> (8 * k)) & 127u);
                int ik = 0;
#pragma unroll
                for (int q = 0; q < RS; ++q) ik = q == k ? idx[q] : ik;
                sts64d(cell_addr(L, k, (p + d) & 31), cell_value(L, k, wrap1(ik + d, T.r[k].n_sigma), 0));
            } else if (HG) {
                const int e2 = e - RS * W0S;
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

    // ---- prologue, all working warps: the chains' current states into shared memory, every window built around position 0
    //      (262 cells per chain at C3: a warp takes every 11th chain) ----
    {
        constexpr int NPW = HG ? 11 : 10;
        const int pw = warp < 4 ? warp : (warp < 8 ? warp - 1 : (warp < 12 ? warp - 2 : warp - 3));
        for (int L = pw; L < 32; L += NPW) {
            long long cL = (long long)blockIdx.x * 32 + L;
            if (cL >= n) cL = n - 1;
            stage_state(L, C.state[cL], C.lam[cL]);
            __syncwarp();
            int idx[R];
#pragma unroll
            for (int k = 0; k < R; ++k) idx[k] = C.isg[(size_t)k * n + cL];
            build_chain(L, 0u, 0u, idx, HG ? C.igm[(size_t)QG * n + cL] : 0);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");
    }

    if ((warp >= 1 && warp <= 3) || warp == 11) {
        // =====================================================================================
        // generators: Philox + decode.  EVERY batch is generated by all four warps, two steps each: what the ring's depth has to
        // cover is the latency of a batch, not the throughput
        // =====================================================================================
        const int g = warp == 11 ? 3 : warp - 1;
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
        for (int b = 0; !it.done(); ++b, it.advance()) {
            const int stage = b % NS, fill = b / NS;
            trace(b, 0);
            if (fill >= 1) { if (TAB_SPIN >= 2) mbar_spin(b_free + 8u * (uint32_t)stage, (uint32_t)(fill - 1) & 1u); else mbar_wait(b_free + 8u * (uint32_t)stage, (uint32_t)(fill - 1) & 1u); }
            trace(b, 1);
            int s0, pad;
            bool snap_last;
            it.get(s0, pad, snap_last);
            const uint32_t rs = ring_of(stage);
            // a ROLLED loop over the steps (the body stays in the instruction cache; the helper warps of an SM run five different
            // code paths): the lower-bound gather of step j is issued in iteration j and consumed in iteration j + 1
            // the batch's proposals as masks (bit 8 r + j: step j proposes a move of restraint r < 4): any nuisance move / a non-null
            // sigma move / its direction; m_g: non-null gamma moves, their directions << 8, state moves << 16, real positions << 24;
            // m_4: the same three masks of restraint 4 in bits 0-7, 8-15, 16-23
            uint32_t m_nm = 0, m_fm = 0, m_up = 0, m_g = 0, m_4 = 0;
            double p_th = inf, p_lb = inf, p_lu = inf;       // step j - 1: threshold, gathered lower bound, ln u - 2 margin
#pragma unroll
            for (int jj = 0; jj < BL / 4; ++jj) {
                const int j = (BL / 4) * g + jj;
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
                const bool nu = real && nuis, sm = real && !nuis;
                // lb + ln u - 2 margin for a state move (the gather is issued for every lane: L2-resident table)
                const double lbv = __ldg(lb + (sm ? hi : 0u));
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
                const bool scal = !HG || hi < (uint32_t)RS;
                const uint32_t cs = (a - 1u) & 127u, cg = (bb - 1u) & 127u;
                const uint32_t w_ds = (nu && scal) ? cs << (8u * (hi & 3u)) : 0u;
                const uint32_t w_dn = (nu && !scal) ? (cs | (cg << 8)) : 0u;
                const uint32_t w_sel = nu ? ((scal ? (0x00107760u | hi) : 0x7754u) | (hi << 16)) : (0x00107760u | (7u << 16));   // bit 20: 32-slot table
                const uint32_t w_off = nu ? (scal ? hi : (uint32_t)RS) * 8192u : 0u;
                const uint32_t w_dw = !real ? TAB_PAD : (nuis ? (0x80000000u | (hi << 4) | (bb << 2) | a) : hi);
                sts128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane, w_ds, w_dn, w_sel, w_off);
                sts32(rs + F::O_DW + 128u * (uint32_t)j + lane4, w_dw);
                {
                    // bit 8 r + j of m_nm / m_fm / m_up: step j proposes a move / a non-null sigma move / an upward one of restraint
                    // r < 4; m_g: non-null gamma proposals, their directions << 8, state-move proposals << 16; m_4: the three
                    // masks of restraint 4 in bits 0-7, 8-15, 16-23
                    const bool lo4 = R <= 4 || hi < 4u;
                    const uint32_t bit = 1u << (8u * (hi & 3u) + (uint32_t)j), bj = 1u << j;
                    m_nm |= (nu && lo4) ? bit : 0u;
                    m_fm |= (nu && lo4 && a != 1u) ? bit : 0u;
                    m_up |= (nu && lo4 && a == 2u) ? bit : 0u;
                    if (R > 4) m_4 |= (nu && !lo4) ? (bj | (a != 1u ? bj << 8 : 0u) | (a == 2u ? bj << 16 : 0u)) : 0u;
                    m_g |= sm ? bj << 16 : 0u;
                    if (HG) m_g |= (nu && hi == (uint32_t)QG) ? ((bb != 1u ? bj : 0u) | (bb == 2u ? bj << 8 : 0u)) : 0u;
                }
                // thresholds of the previous step (its gather has had an iteration to land), then this step's become pending
                if (jj > 0) sts128(rs + F::O_TX + 512u * (uint32_t)(j - 1) + 16u * (uint32_t)lane, make_double2(p_th, __dadd_rn(p_lb, p_lu)));
                p_th = nu ? __dadd_rn(lud, TAB_MARGIN) : inf;
                p_lb = lbv;
                p_lu = sm ? __dsub_rn(lud, 2.0 * TAB_MARGIN) : inf;
            }
            sts128(rs + F::O_TX + 512u * (uint32_t)((BL / 4) * g + BL / 4 - 1) + 16u * (uint32_t)lane, make_double2(p_th, __dadd_rn(p_lb, p_lu)));
            // (the masks of this warp's two steps; the accountants OR the four parts)
            sts128u(rs + F::O_MK + 640u * (uint32_t)g + 16u * (uint32_t)lane, m_nm, m_fm, m_up, m_g);
            if (R > 4) sts32(rs + F::O_MK + 640u * (uint32_t)g + 512u + lane4, m_4);
            trace(b, 2);
            mbar_arrive(b_full + 8u * (uint32_t)stage);
        }
        return;
    }

    if (warp == 10) {
        // =====================================================================================
        // tabulator: keeps every chain's windows around its walkers
        // =====================================================================================
        int nf[NF > 0 ? NF : 1];                             // grid length of every field
        int pfull[NF > 0 ? NF : 1], ifull[NF > 0 ? NF : 1], lo[NF > 0 ? NF : 1], hi[NF > 0 ? NF : 1];
        int lop[NF > 0 ? NF : 1], hip[NF > 0 ? NF : 1], lo_cur[NF > 0 ? NF : 1], hi_cur[NF > 0 ? NF : 1];   // the window one batch ago (and a copy)
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int k = f < R ? f : QG;
            nf[f] = f < R ? T.r[k].n_sigma : G;
            ifull[f] = f < R ? C.isg[(size_t)k * n + c] : C.igm[(size_t)QG * n + c];
            pfull[f] = 0; lo[f] = f < RS ? -TAB_MRG0S : -TAB_MRG0; hi[f] = -lo[f];
            lop[f] = lo[f]; hip[f] = hi[f]; lo_cur[f] = lo[f]; hi_cur[f] = hi[f];
        }
        const uint32_t lane_lt = (1u << lane) - 1u;
        uint32_t n1 = 0, n2 = 0;
        // fill queued cells: scalars one cell per item, the gamma restraint one row / column segment per item
        auto process1 = [&]() {
            __syncwarp();
            for (uint32_t i = (uint32_t)lane; i < n1; i += 32) {
                const uint32_t w = lds32(q10 + 4u * i);
                const int L = (int)(w & 31u), k = (int)((w >> 5) & 7u), slot = (int)((w >> 8) & 31u), ii = (int)(w >> 13);
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

        constexpr uint32_t MS = 0x7f7f7f7fu, MN = 0x00007f7fu, HS = 0x80808080u, HN = 0x00008080u;
        // packed per-field words: the published windows {-lo, 127 - width} and the trigger windows [lo + MRGMIN, hi - MRGMIN]
        uint32_t nls, mgs, nln, mgn, tls, tgs, tln, tgn;
        auto pack = [&]() {
            nls = 0; mgs = 0; nln = 0; mgn = 0; tls = 0; tgs = 0; tln = 0; tgn = 0;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const uint32_t nl = (uint32_t)(-lo[f]) & 127u, mg = (uint32_t)(127 - (hi[f] - lo[f]));
                const uint32_t tl = (uint32_t)(-(lo[f] + TAB_MRGMIN)) & 127u, tg = (uint32_t)(127 - (hi[f] - lo[f] - 2 * TAB_MRGMIN));
                if (f < RS) { nls |= nl << (8 * f); mgs |= mg << (8 * f); tls |= tl << (8 * f); tgs |= tg << (8 * f); }
                else { nln |= nl << (8 * (f - RS)); mgn |= mg << (8 * (f - RS)); tln |= tl << (8 * (f - RS)); tgn |= tg << (8 * (f - RS)); }
            }
#pragma unroll
            for (int f = RS; f < 4; ++f) { mgs |= 127u << (8 * f); tgs |= 127u << (8 * f); }     // unused fields: position 0 in window [0, 0]
            if (!HG) { mgn = 127u | (127u << 8); tgn = mgn; }
        };
        pack();
        uint32_t ckps = 0u, ckpn = 0u;                       // the positions pfull / ifull belong to (a checkpoint, refreshed below)

        for (int b = 0; !it.done(); ++b, it.advance()) {
            const int stage = b % NS, fill = b / NS;
            const uint32_t rs = ring_of(stage);
            trace(b, 0);
            if (TAB_SPIN >= 2) mbar_spin(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u); else mbar_wait(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
            trace(b, 1);
            const uint32_t ps = lds32(rs + F::O_PS + lane4), pn = lds32(rs + F::O_PN + lane4);
            const bool reb = (lds32(rs + F::O_ACC + lane4) >> 16) & 1u;
            mbar_arrive(b_free + 8u * (uint32_t)stage);
            if (TAB_T_NOTAB) { mbar_arrive(b_tab + 8u * (uint32_t)(b % 3)); continue; }
            // walkers that left their trigger window (packed test, like the consumer's), chains rebuilt by the careful path
            // (per scalar restraint: bit 8 f + 7 of the warp's OR -- the window policy and the cell loops skip the others)
            const uint32_t any_fs = RS > 0 ? __reduce_or_sync(SPEC_FULL, reb ? HS : (((ps + tls) & MS) + tgs) & HS) : 0u;
            const bool trig_n = HG && (reb || ((((pn + tln) & MN) + tgn) & HN) != 0u);
            const bool any_s = any_fs != 0u, any_n = HG && __any_sync(SPEC_FULL, trig_n);
            // the scalar restraints and the gamma restraint are handled separately: with 32-slot windows the scalars rarely trigger
            trace_t(b, 0);
            const bool per = (b & 3) == 3, do_s = any_s || (RS > 0 && per), do_n = any_n || (HG && per);
            if (do_s || do_n) {
                // checkpoint: unwrapped positions and true indices at the end of this batch (at most 4 batches = 32 steps on)
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    if (!(f < RS ? do_s : do_n)) continue;
                    const uint32_t pm = f < RS ? (ps >> (8 * f)) & 127u : (pn >> (8 * (f - RS))) & 127u;
                    const int d = (int)((pm - (uint32_t)pfull[f] + 64u) & 127u) - 64;
                    pfull[f] += d;
                    ifull[f] = wrap1(ifull[f] + d, nf[f]);
                }
                if (do_s) ckps = ps;
                if (do_n) ckpn = pn;
            }
            trace_t(b, 1);
            if (any_s || any_n) {
                int lo_n[NF > 0 ? NF : 1], hi_n[NF > 0 ? NF : 1];
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    lo_n[f] = lo[f]; hi_n[f] = hi[f];
                    if (!(f < RS ? (any_fs >> (8 * f + 7)) & 1u : (uint32_t)any_n)) continue;
                    const int span = f < RS ? 31 : 15, mrg = f < RS ? TAB_MRG_S : TAB_MRG_N;     // slots - 1
                    const int p = pfull[f];
                    int l = lo[f], h = hi[f];
                    if (reb) { l = p - (f < RS ? TAB_MRG0S : TAB_MRG0); h = p + (f < RS ? TAB_MRG0S : TAB_MRG0); lo[f] = l; hi[f] = h; lop[f] = l; hip[f] = h; }   // rebuilt by the consumer's careful path
                    else {
                        // extend towards the walker; the new window and the two before it together span at most `slots` positions, so
                        // a slot that the consumer may still read (it works with one of the two previous windows) is never rewritten
                        if (hi[f] - p < TAB_MRGMIN) h = imax(hi[f], imin(p + mrg, imin(lo[f], lop[f]) + span));
                        if (p - lo[f] < TAB_MRGMIN) l = imin(lo[f], imax(p - mrg, imax(hi[f], hip[f]) - span));
                        if (h - l > span - 1) { if (p - l > h - p) l = h - (span - 1); else h = l + (span - 1); }
                    }
                    lo_n[f] = l; hi_n[f] = h;
                }
                trace_t(b, 2);
                // ---- new cells: scalars ----
                if (any_s) {
#pragma unroll
                    for (int f = 0; f < RS; ++f) {
                        if (!((any_fs >> (8 * f + 7)) & 1u)) continue;
                        // positions hi+1 .. hi_n, then lo_n .. lo-1 (either range may be empty)
                        const int up0 = imax(hi[f] + 1, lo_n[f]), nup = imax(hi_n[f] - up0 + 1, 0);
                        const int dn1 = imin(lo[f] - 1, hi_n[f]), ndn = imax(dn1 - lo_n[f] + 1, 0);
                        const int cnt = nup + ndn;
#if TAB_CELLPAR
                        // one (chain, restraint) at a time, its new cells (at most 31) across the lanes: few chains move a window
                        // in a batch, and a loop over cells with a vote per cell is what a single warp is slowest at
                        uint32_t m = __ballot_sync(SPEC_FULL, cnt > 0);
                        while (m) {
                            const int L = __ffs((int)m) - 1;
                            m &= m - 1;
                            const int up0L = __shfl_sync(SPEC_FULL, up0, L), nupL = __shfl_sync(SPEC_FULL, nup, L);
                            const int loL = __shfl_sync(SPEC_FULL, lo_n[f], L), cntL = __shfl_sync(SPEC_FULL, cnt, L);
                            const int baseL = __shfl_sync(SPEC_FULL, ifull[f] - pfull[f], L);
                            if (lane < cntL) {
                                const int q = lane < nupL ? up0L + lane : loL + (lane - nupL);
                                sts64d(cell_addr(L, f, q & 31), cell_value(L, f, wrap1(baseL + q, T.r[f].n_sigma), 0));
                            }
                        }
#else
                        for (int i = 0; __any_sync(SPEC_FULL, i < cnt); ++i) {
                            const int q = i < nup ? up0 + i : lo_n[f] + (i - nup);
                            const int ii = wrap1(ifull[f] + (q - pfull[f]), nf[f]);
                            emit1(i < cnt, (uint32_t)lane | ((uint32_t)f << 5) | ((uint32_t)(q & 31) << 8) | ((uint32_t)ii << 13));
                        }
#endif
                    }
                    if (!TAB_CELLPAR && n1) process1();
                    __syncwarp();
                }
                trace(b, 3);
                // ---- new cells: the gamma restraint (new sigma columns over the new gamma range, new gamma rows over the new
                //      sigma range; cells in both are filled twice with the same value) ----
                if (HG && any_n) {
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
                    trace_t(b, 4);
                    if (n2) process2();
                }
                trace_t(b, 5);
#pragma unroll
                for (int f = 0; f < NF; ++f) { lo[f] = lo_n[f]; hi[f] = hi_n[f]; }
                pack();
                trace_t(b, 6);
            }
            // (the window of the batch before: a batch that changes nothing ages the older one out as well)
#pragma unroll
            for (int f = 0; f < NF; ++f) { lop[f] = lo_cur[f]; hip[f] = hi_cur[f]; lo_cur[f] = lo[f]; hi_cur[f] = hi[f]; }
            // ---- publish: the windows (valid two batches on), the checkpoint {positions, true indices} ----
            const uint32_t rb = rect0 + (uint32_t)(b % 3) * 768u + lane4;
            sts32(rb, nls); sts32(rb + 128u, mgs); sts32(rb + 256u, nln); sts32(rb + 384u, mgn);
            sts32(rb + 512u, ckps); sts32(rb + 640u, ckpn);
            const uint32_t ib = idxp0 + (uint32_t)(b % 3) * (uint32_t)(R + 1) * 128u + lane4;
#pragma unroll
            for (int k = 0; k < R; ++k) sts32(ib + 128u * (uint32_t)k, (uint32_t)ifull[k]);
            if (HG) sts32(ib + 128u * (uint32_t)R, (uint32_t)ifull[NF - 1]);
            trace(b, 2);
            mbar_arrive(b_tab + 8u * (uint32_t)(b % 3));
        }
        return;
    }

    if (warp == 5 || warp == 6 || warp == 7 || warp == 9 || warp == 13) {
        // =====================================================================================
        // accountants: the bookkeeping of the batches the consumer has finished, split by PARAMETER over four warps (a
        // single warp runs dependent integer code at ~0.2 instructions per cycle; the parameters are independent).  ONE code
        // path for the four (the kernel's hot code has to stay small: the instruction caches of 128 SMs refill from L2):
        // accountant a < 4 follows the sigma index of restraint a in slot 0 (accountant 2 that of restraint 4 in slot 1), accountant
        // 4 (warp 13) the gamma index; accountant 0 also keeps the conformational state, the state trace and the acceptance totals.
        // An index and the start of its current run follow from the generator's proposal masks AND the accept bits.
        // =====================================================================================
        const int aw = warp == 9 ? 3 : (warp == 13 ? 4 : warp - 5);
        const bool has0 = aw < R && aw < 4, has1 = (aw == 4 && HG) || (aw == 2 && R > 4);
        const int q0 = has0 ? aw : 0, q1 = aw == 4 ? R : 4;               // parameter: restraint's sigma, or R = gamma
        const int grp = C.group[c];
        unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
        unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
        const bool tracing = aw == 0 && A.state_trace != nullptr;
        const bool traced = tracing && live && c < A.trace_n;
        // a window of 16 histogram bins per chain and parameter in shared memory ([parameter][bin][chain], private to the
        // lane: plain read-modify-write): a chain keeps revisiting the same few bins
        const uint32_t hwdummy = acc0 + (uint32_t)(R + 1) * 16u * 128u + lane4;
        struct Slot { int idx, since, wb, nn; uint32_t hw; unsigned long long* bins; unsigned int acc; int pos, wlo; };
        Slot sl[2];
        sl[0].nn = T.r[q0].n_sigma; sl[0].bins = g_param_counts + T.r[q0].hist_sigma; sl[0].idx = C.isg[(size_t)q0 * n + c];
        sl[1].nn = q1 == R ? G : T.r[R > 4 ? 4 : 0].n_sigma;
        sl[1].bins = g_param_counts + (q1 == R ? T.r[QG].hist_gamma : T.r[R > 4 ? 4 : 0].hist_sigma);
        sl[1].idx = q1 == R ? (HG ? C.igm[(size_t)QG * n + c] : 0) : C.isg[(size_t)(R > 4 ? 4 : 0) * n + c];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            sl[u].since = burn; sl[u].acc = 0;
            sl[u].hw = acc0 + (uint32_t)((u ? q1 : q0) * 16) * 128u + lane4;
            sl[u].wb = sl[u].idx - 8; sl[u].wb += sl[u].wb < 0 ? sl[u].nn : 0;
            sl[u].pos = 0; sl[u].wlo = -8;
            if (u ? has1 : has0)
                for (int bq = 0; bq < 16; ++bq) sts32(sl[u].hw + (uint32_t)bq * 128u, 0u);
        }
        int state = C.state[c], since_state = burn;
        unsigned int acc_state = 0, acc_total = 0;
        // the events of one parameter in a batch, in step order (most lanes have none or one): {bin, run length} into its
        // window; a walker that has left the window flushes it first
        // one event of a parameter: {bin, run length} into its window; a walker that has left the window flushes it first
        auto event1 = [&](Slot& z, uint32_t& m, uint32_t ups, int s0, int pad) {
            const int j = __ffs((int)m) - 1;                                 // -1 where the lane has no event left
            const int s = s0 + j - pad;
            const bool on = m != 0u, rec_on = s > burn;
            const bool ev = on && rec_on && live && s > z.since;
#if TAB_SLIDE
            // a SLIDING window of 16 bins, slot = position & 15, that always holds the walker: the run goes to the walker's bin, then
            // the walker moves; a move that pushes the window's edge first empties the one slot the walker is about to take
            {
                const uint32_t a = ev ? z.hw + (uint32_t)(z.pos & 15) * 128u : hwdummy;
                sts32(a, lds32(a) + (uint32_t)(s - z.since));
                if (on) {
                    const int d = ((ups >> j) & 1u) ? 1 : -1;
                    z.idx = wrap1(z.idx + d, z.nn);
                    z.pos += d;
                    if ((uint32_t)(z.pos - z.wlo) >= 16u) {
                        const uint32_t av = z.hw + (uint32_t)(z.pos & 15) * 128u;
                        const uint32_t v = lds32(av);
                        if (v) {
                            int gi = z.idx - 16 * d;                  // the grid index of the position that leaves the window
                            gi += gi < 0 ? z.nn : 0;
                            gi -= gi >= z.nn ? z.nn : 0;
                            atomicAdd(z.bins + gi, (unsigned long long)v);
                            sts32(av, 0u);
                        }
                        z.wlo += d;
                    }
                }
                z.since = (on && rec_on) ? s : z.since;
                m &= m - 1u;
                return;
            }
#endif
            int off = z.idx - z.wb;
            off += off < 0 ? z.nn : 0;
            if (ev && (uint32_t)off >= 16u) {                                // (rare, divergent: no vote)
                flush_window(z.hw, z.wb, z.nn, z.bins);
                z.wb = z.idx - 8;
                z.wb += z.wb < 0 ? z.nn : 0;
                off = 8;
            }
            const uint32_t a = ev ? z.hw + (uint32_t)off * 128u : hwdummy;
            sts32(a, lds32(a) + (uint32_t)(s - z.since));
            z.idx = on ? wrap1(z.idx + (((ups >> j) & 1u) ? 1 : -1), z.nn) : z.idx;
            z.since = (on && rec_on) ? s : z.since;
            m &= m - 1u;
        };
        // the events of one or two parameters in a batch, in step order per parameter, the two side by side (most lanes have none
        // or one event per parameter and batch)
        auto events = [&](Slot& z, uint32_t m, uint32_t ups, int s0, int pad) {
            while (__any_sync(SPEC_FULL, m != 0u)) event1(z, m, ups, s0, pad);
        };
        auto events2 = [&](Slot& za, uint32_t ma, uint32_t ua, Slot& zb, uint32_t mb, uint32_t ub, int s0, int pad) {
            while (__any_sync(SPEC_FULL, (ma | mb) != 0u)) {
                event1(za, ma, ua, s0, pad);
                event1(zb, mb, ub, s0, pad);
            }
        };

        for (int b = 0; !it.done(); ++b, it.advance()) {
            const int stage = b % NS, fill = b / NS;
            const uint32_t rs = ring_of(stage);
            int s0, pad;
            bool snap_last;
            it.get(s0, pad, snap_last);
            trace(b, 0);
            if (TAB_SPIN >= 2) mbar_spin(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u); else mbar_wait(b_empty + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
            trace(b, 1);
            uint4 mk = lds128u(rs + F::O_MK + 16u * (uint32_t)lane);
            uint32_t mk4 = R > 4 ? lds32(rs + F::O_MK + 512u + lane4) : 0u;
#pragma unroll
            for (int gq = 1; gq < 4; ++gq) {
                const uint4 v = lds128u(rs + F::O_MK + 640u * (uint32_t)gq + 16u * (uint32_t)lane);
                mk.x |= v.x; mk.y |= v.y; mk.z |= v.z; mk.w |= v.w;
                if (R > 4) mk4 |= lds32(rs + F::O_MK + 640u * (uint32_t)gq + 512u + lane4);
            }
            const uint32_t accb = lds32(rs + F::O_ACC + lane4) & 0xffu;
            const double E_end = lds64(rs + F::O_E + lane8);
            // accepted state moves (rare on large ensembles), the optional state trace: the raw proposal words are needed
            const uint32_t stm = aw == 0 ? (mk.w >> 16) & accb : 0u;
            const bool need_dw = aw == 0 && (tracing || __any_sync(SPEC_FULL, stm != 0u));
            uint32_t w_dw[BL];
            if (need_dw) {
#pragma unroll
                for (int j = 0; j < BL; ++j) w_dw[j] = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
            }
            if (!TAB_T_AFREE) mbar_arrive(b_free + 8u * (uint32_t)stage);
            if (TAB_T_NOACC) continue;
            const uint32_t accrep = accb * 0x01010101u;
            if (aw == 0) acc_total += (unsigned int)__popc(accb);
            uint32_t m0 = 0, u0 = 0, m1 = 0, u1 = 0;
            if (has0) {
                sl[0].acc += (unsigned int)__popc(((mk.x & accrep) >> (8 * q0)) & 0xffu);
                m0 = ((mk.y & accrep) >> (8 * q0)) & 0xffu; u0 = (mk.z >> (8 * q0)) & 0xffu;
            }
            if (has1) {
                if (q1 == R) { m1 = mk.w & accb; u1 = (mk.w >> 8) & 0xffu; }
                else {
                    sl[1].acc += (unsigned int)__popc(mk4 & accb);
                    m1 = (mk4 >> 8) & accb; u1 = (mk4 >> 16) & 0xffu;
                }
            }
            if (has1) events2(sl[0], m0, u0, sl[1], m1, u1, s0, pad);
            else events(sl[0], m0, u0, s0, pad);
            // accepted state moves and the optional state trace, in step order
            if (need_dw) {
#pragma unroll
                for (int j = 0; j < BL; ++j) {
                    const uint32_t dw = w_dw[j];
                    const int s = s0 + j - pad;
                    if ((stm >> j) & 1u) {
                        ++acc_state;
                        const int i2 = (int)dw;
                        if (i2 != state) {
                            const bool rec_on = s > burn;
                            if (rec_on && live) atomicAdd(g_state_counts + state, (unsigned long long)(s - since_state));
                            if (rec_on) since_state = s;
                            state = i2;
                        }
                    }
                    if (traced && dw != TAB_PAD && s >= burn) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
                }
            }
            trace(b, 2);
            if (snap_last) {
                if (live) {
                    const size_t o = (size_t)snap_idx * n + c;
                    if (aw == 0) {
                        A.traj_E[o] = E_end;
                        A.traj_state[o] = state;
                        A.traj_accept[o] = (unsigned char)((accb >> (BL - 1)) & 1u);
                    }
                    if (has0) A.traj_idx[((size_t)snap_idx * T.P + q0) * n + c] = sl[0].idx;
                    if (has1) A.traj_idx[((size_t)snap_idx * T.P + q1) * n + c] = sl[1].idx;
                }
                ++snap_idx;
            }
        }
        // ---- flush the run-length counters and the windows, persist the own part of the chain (E: the consumer) ----
        if (live) {
            if (aw == 0) {
                if (s_end > since_state) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - since_state));
                C.state[c] = state;
                C.accepted[c] += acc_total;
                C.sep[(size_t)R * n + c] += acc_state;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!(u ? has1 : has0)) continue;
                Slot& z = sl[u];
                const int q = u ? q1 : q0;
                if (s_end > z.since) atomicAdd(z.bins + z.idx, (unsigned long long)(s_end - z.since));
                for (int bq = 0; bq < 16; ++bq) {
                    const uint32_t v = lds32(z.hw + (uint32_t)bq * 128u);
                    int gi = z.wb + bq;
                    gi -= gi >= z.nn ? z.nn : 0;
                    if (TAB_SLIDE) {                                 // slot bq holds the window's position p with p & 15 == bq
                        const int pp = z.wlo + ((bq - z.wlo) & 15);
                        gi = z.idx + (pp - z.pos);
                        gi += gi < 0 ? z.nn : 0;
                        gi -= gi >= z.nn ? z.nn : 0;
                    }
                    if (v) atomicAdd(z.bins + gi, (unsigned long long)v);
                }
                if (q == R) C.igm[(size_t)QG * n + c] = z.idx;
                else { C.isg[(size_t)q * n + c] = z.idx; C.sep[(size_t)q * n + c] += z.acc; }
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
    int cstate = C.state[c];                                 // the chain's conformational state (changes in the careful path only)
    unsigned int n_batches = 0, n_rollbacks = 0;
    const uint32_t tabl = tab0 + lane8;
    constexpr uint32_t MS = 0x7f7f7f7fu, MN = 0x00007f7fu;
    constexpr uint32_t HS = 0x80808080u, HN = 0x00008080u;
    // windows: valid where ((P + neglo) & mask) + marg has no bit 7 set in any field
    constexpr uint32_t NL0 = (uint32_t)TAB_MRG0 * 0x01010101u, MG0 = (uint32_t)(127 - 2 * TAB_MRG0) * 0x01010101u;
    constexpr uint32_t NL0S = (uint32_t)TAB_MRG0S * 0x01010101u, MG0S = (uint32_t)(127 - 2 * TAB_MRG0S) * 0x01010101u;
    uint32_t ov_nls = 0, ov_mgs = 0, ov_nln = 0, ov_mgn = 0;  // windows of the chains the careful path has just rebuilt
    int rebat = -1;                                           // the batch of the chain's last rebuild

    // the Metropolis uniform of local step s (the rare paths only: the ring does not carry it)
    auto uniform_of = [&](int s) {
        const unsigned long long ctr = A.step0 + (unsigned long long)A.s_begin + (unsigned long long)(long long)s;
        const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
        if (TAB_NI_U) return step_uniform(ctr, chain_id, C.seed);
        const Philox4 w = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)chain_id, (uint32_t)(chain_id >> 32),
                                        (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
        return spec_u2(w.w2, w.w3);
    };
    auto metropolis_accept = [&](double d, double u) { return TAB_NI_A ? accept_exact(d, u) : bcp::metropolis_accept(d, u); };
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

    // true indices at packed positions (psj, pnj): the tabulator's checkpoint of parity buffer `par` (the launch's start if
    // !have) plus the positions' progress since (at most 63 positions: checkpoints are at most 6 batches old)
    auto true_indices = [&](bool have, int par, uint32_t psj, uint32_t pnj, int* idj, int& gj) {
        uint32_t psb = 0u, pnb = 0u;
        int gib = gi0;
#pragma unroll
        for (int k = 0; k < R; ++k) idj[k] = idx0[k];
        if (have) {
            const uint32_t rb = rect0 + (uint32_t)par * 768u + lane4;
            const uint32_t ib = idxp0 + (uint32_t)par * (uint32_t)(R + 1) * 128u + lane4;
            psb = lds32(rb + 512u); pnb = lds32(rb + 640u);
#pragma unroll
            for (int k = 0; k < R; ++k) idj[k] = (int)lds32(ib + 128u * (uint32_t)k);
            if (HG) gib = (int)lds32(ib + 128u * (uint32_t)R);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const uint32_t pm = (HG && k == QG) ? (pnj & 127u) - (pnb & 127u) : ((psj >> (8 * k)) & 127u) - ((psb >> (8 * k)) & 127u);
            idj[k] = wrap1(idj[k] + ((int)((pm + 64u) & 127u) - 64), T.r[k].n_sigma);
        }
        gj = HG ? wrap1(gib + ((int)(((((pnj >> 8) & 127u) - ((pnb >> 8) & 127u)) + 64u) & 127u) - 64), G) : 0;
    };

#pragma unroll
    for (int k = 0; k < R; ++k) t[k] = lds64(tabl + (uint32_t)((HG && k == QG) ? 32 * RS : 32 * k) * 256u);

    bool ok_full = false, ok_tab = false;                     // the batch's barriers have been seen completed already
    for (int b = 0; !it.done(); ++b, it.advance()) {
        const int stage = b % NS, fill = b / NS;
        const uint32_t rs = ring_of(stage);
        trace(b, 0);
        if (!ok_full) mbar_wait(b_full + 8u * (uint32_t)stage, (uint32_t)fill & 1u);
        trace(b, 1);
        int cs0, cpad;                                        // the batch's first real step and front padding (rare paths)
        {
            bool snap_last;
            it.get(cs0, cpad, snap_last);
        }
        uint32_t nls, mgs, nln, mgn;
        int xw = -1;                                          // the tabulator batch whose windows / checkpoint this batch works with
        if (b >= 2) {
            // the windows after batch b - 2 if the tabulator has got that far, else those after batch b - 3 (it never rewrites a
            // slot of the two windows before the one it is building)
            xw = b - 2;
            if (!ok_tab && !mbar_test(b_tab + 8u * (uint32_t)(xw % 3), (uint32_t)(xw / 3) & 1u)) {
                if (b >= 3) --xw;
                mbar_wait(b_tab + 8u * (uint32_t)(xw % 3), (uint32_t)(xw / 3) & 1u);
            }
            const uint32_t rb = rect0 + (uint32_t)(xw % 3) * 768u + lane4;
            nls = lds32(rb); mgs = lds32(rb + 128u); nln = lds32(rb + 256u); mgn = lds32(rb + 384u);
        } else {
            nls = NL0S; mgs = MG0S; nln = NL0 & MN; mgn = MG0 & 0xffffu;
#pragma unroll
            for (int f = RS; f < 4; ++f) { nls &= ~(0xffu << (8 * f)); mgs |= 127u << (8 * f); }
            if (!HG) { nln = 0u; mgn = 127u | (127u << 8); }
        }
        // a chain the careful path has rebuilt keeps its own window until the tabulator's windows include that batch
        if (xw < rebat) { nls = ov_nls; mgs = ov_mgs; nln = ov_nln; mgn = ov_mgn; }
        trace(b, 2);

        // ---- fast batch ----------------------------------------------------------------------------
        const double sv_E = E;
        const uint32_t sv_Ps = Ps, sv_Pn = Pn;
        double sv_t[R];
#pragma unroll
        for (int k = 0; k < R; ++k) sv_t[k] = t[k];
        uint32_t accb = 0, why = 0;
        bool rare = false;
        // the lane's first state-move proposal of this batch that the lower bound cannot reject: step, energy and positions then
        int jf = -1;
        double Ef = 0.0;
        uint32_t Psf = 0u, Pnf = 0u;
#if TAB_SPEC2
        // One-step speculation on the table load: while step j is resolved, the proposal of step j + 1 is looked up under both
        // outcomes of step j (positions P + D and P' + D: the deltas come from the counter-based RNG, not from the chain), so the
        // address computation and the shared-memory load leave the dependent chain accept -> select -> 3 adds -> compare.
        auto lookup = [&](uint32_t ps2, uint32_t pn2, const uint4& ww, double& tv, bool& bd) {
            const uint32_t x = __byte_perm(ps2, pn2, ww.z);
            const uint32_t e = (x & (0xfu | ((ww.z >> 16) & 0x10u))) | ((x >> 4) & 0xf0u);
            tv = lds64(tabl + ww.w + e * 256u);
            const uint32_t rr = (ww.z >> 16) & 7u;
            const uint32_t bs = (((ps2 + nls) & MS) + mgs) & HS, bn = (((pn2 + nln) & MN) + mgn) & HN;
            bd = !TAB_T_NOTAB && (((ww.z >> 20) & 1u) ? (bs >> (8 * (rr & 3))) & 0x80u : bn) != 0u;
        };
        uint4 w = lds128u(rs + F::O_W + 16u * (uint32_t)lane);
        uint4 w1 = lds128u(rs + F::O_W + 512u + 16u * (uint32_t)lane);
        double2 txv = lds128(rs + F::O_TX + 16u * (uint32_t)lane);
        uint32_t Ps2 = (Ps + w.x) & MS, Pn2 = (Pn + w.y) & MN;         // the pending proposal (step 0: looked up directly)
        double tp;
        bool bad;
        lookup(Ps2, Pn2, w, tp, bad);
#pragma unroll 1
        for (int j = 0; j < BL; ++j) {
            // words two steps ahead (the last iterations re-read the batch's last ones: the speculation of the step after the
            // batch's last is thrown away)
            const uint32_t j2 = (uint32_t)(j + 2 < BL ? j + 2 : BL - 1), j1 = (uint32_t)(j + 1 < BL ? j + 1 : BL - 1);
            const uint4 w2 = lds128u(rs + F::O_W + 512u * j2 + 16u * (uint32_t)lane);
            const double2 txn = lds128(rs + F::O_TX + 512u * j1 + 16u * (uint32_t)lane);
            // step j + 1 under both outcomes of step j
            const uint32_t PAs = (Ps + w1.x) & MS, PAn = (Pn + w1.y) & MN, PBs = (Ps2 + w1.x) & MS, PBn = (Pn2 + w1.y) & MN;
            double tA, tB;
            bool bA, bB;
            lookup(PAs, PAn, w1, tA, bA);
            lookup(PBs, PBn, w1, tB, bB);
            // resolve step j
            const int r = (int)((w.z >> 16) & 7u);
            double xk[R];
#pragma unroll
            for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t[k];
            double En = xk[0];
#pragma unroll
            for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
            const double Ahi = __dsub_rn(E, txv.x), Alo = __dadd_rn(Ahi, 2.0 * TAB_MARGIN);
            const bool acc = En < Ahi;
            const bool unsure = !acc && !(En > Alo);                     // guard band, also NaN
            // a state move the lower bound cannot reject: the first one of a lane is resolved after the batch, a second one rolls back
            const bool lbf = !(E < txv.y), first = lbf && jf < 0;
            rare = rare || unsure || (lbf && (!first || (TAB_DBG & 4))) || bad;
            if (TAB_STATS) { why |= bad ? 2u : 0u; why |= unsure ? 1u : 0u; }
            Ef = first ? E : Ef; Psf = first ? Ps : Psf; Pnf = first ? Pn : Pnf; jf = first ? j : jf;
            E = acc ? En : E;
#pragma unroll
            for (int k = 0; k < R; ++k) t[k] = acc ? xk[k] : t[k];
            Ps = acc ? Ps2 : Ps;
            Pn = acc ? Pn2 : Pn;
            accb |= acc ? (1u << j) : 0u;
            // the pending proposal of step j + 1
            Ps2 = acc ? PBs : PAs; Pn2 = acc ? PBn : PAn; tp = acc ? tB : tA; bad = acc ? bB : bA;
            w = w1; w1 = w2; txv = txn;
        }
#else
        uint4 w = lds128u(rs + F::O_W + 16u * (uint32_t)lane);
        double2 txv = lds128(rs + F::O_TX + 16u * (uint32_t)lane);
        constexpr int CU = TAB_CU;
#pragma unroll CU
        for (int j = 0; j < BL; ++j) {
            // (the next step's words are loaded a step ahead; the last iteration reads the 512 bytes after its array -- the next
            // array of the stage -- and drops them)
            const uint4 wn = lds128u(rs + F::O_W + 512u * (uint32_t)(j + 1) + 16u * (uint32_t)lane);
            const double2 txn = lds128(rs + F::O_TX + 512u * (uint32_t)(j + 1) + 16u * (uint32_t)lane);
            const uint32_t Ps2 = (Ps + w.x) & MS, Pn2 = (Pn + w.y) & MN;
            const uint32_t x = __byte_perm(Ps2, Pn2, w.z);
            const uint32_t e = (x & (0xfu | ((w.z >> 16) & 0x10u))) | ((x >> 4) & 0xf0u);
            double tp = lds64(tabl + w.w + e * 256u);
            const int r = (int)((w.z >> 16) & 7u);
            // the proposal's position outside its tabulated window (a walker that outran the tabulator): the term is evaluated
            // here, exactly like a table entry -- rare, and the batch goes on
            // (branch-free: the scalar word's flag of restraint r, or the gamma restraint's two flags)
            const uint32_t bs = (((Ps2 + nls) & MS) + mgs) & HS, bn = (((Pn2 + nln) & MN) + mgn) & HN;
            const uint32_t scal = (w.z >> 20) & 1u;
            const uint32_t bsel = (scal ? (bs >> (8 * (r & 3))) & 0x80u : bn) | (((TAB_DBG & 16) && scal && r < R) ? 1u : 0u) | (((TAB_DBG & 8) && !scal) ? 1u : 0u);
            const bool bad = !TAB_T_NOTAB && bsel != 0u;
            if (TAB_INPLACE && __any_sync(SPEC_FULL, bad)) {
                if (bad) {
                    int idj[R], gj, ir = 0;
                    true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Ps2, Pn2, idj, gj);
#pragma unroll
                    for (int k = 0; k < R; ++k) ir = r == k ? idj[k] : ir;
                    tp = cell_value(lane, r, ir, gj);
                }
                if (TAB_STATS) why |= bad ? 2u : 0u;
            }
            double xk[R];
#pragma unroll
            for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t[k];
            double En = xk[0];
#pragma unroll
            for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
            const double Ahi = __dsub_rn(E, txv.x), Alo = __dadd_rn(Ahi, 2.0 * TAB_MARGIN);
            const bool acc = En < Ahi;
            const bool unsure = !acc && !(En > Alo);                     // guard band, also NaN
            // a state move the lower bound cannot reject: the first one of a lane is resolved after the batch, a second one rolls back
            const bool lbf = !(E < txv.y), first = lbf && jf < 0;
            rare = rare || unsure || (lbf && (!first || (TAB_DBG & 4))) || (!TAB_INPLACE && bad);
            if (TAB_STATS && !TAB_INPLACE) why |= bad ? 2u : 0u;
            Ef = first ? E : Ef; Psf = first ? Ps : Psf; Pnf = first ? Pn : Pnf; jf = first ? j : jf;
            if (TAB_STATS) why |= unsure ? 1u : 0u;
            E = acc ? En : E;
#pragma unroll
            for (int k = 0; k < R; ++k) t[k] = acc ? xk[k] : t[k];
            Ps = acc ? Ps2 : Ps;
            Pn = acc ? Pn2 : Pn;
            accb |= acc ? (1u << j) : 0u;
            w = wn; txv = txn;
        }
#endif
        ++n_batches;
        // the next batch's barriers are tested here, a barrier query's latency ahead of their use (a phase that has completed
        // stays completed until this warp has moved on: the next batch then starts without a wait)
        {
            const int bn = b + 1, stn = bn % NS, fn = bn / NS;
            ok_full = mbar_test(b_full + 8u * (uint32_t)stn, (uint32_t)fn & 1u);
            ok_tab = bn >= 2 ? mbar_test(b_tab + 8u * (uint32_t)((bn - 2) % 3), (uint32_t)((bn - 2) / 3) & 1u) : true;
        }
        uint32_t rebw = 0;
        // (one vote decides whether anything at all needs a second look: almost always nothing does)
        const bool second_look = __any_sync(SPEC_FULL, rare || jf >= 0);
        const long long t_sl0 = TAB_STATS ? clock64() : 0ll;
        bool did_careful = false;
        bool reb = false;                                     // the lane's chain has changed its state in this batch: its tables are stale
        if (second_look && !__any_sync(SPEC_FULL, rare)) {
            // ---- state-move proposals the lower bound could not reject: evaluated exactly at the chain's indices of that step
            //      (true indices = the tabulator's last checkpoint + the positions' progress since); almost all are rejected
            //      and the batch stands ----
            bool moved = false;                               // ... accepted, to another state
            int st_new = cstate;
            double E_mv = 0.0, t_mv[R];
#pragma unroll
            for (int k = 0; k < R; ++k) t_mv[k] = 0.0;
            if (jf >= 0) {
                int idj[R], gj;
                true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Psf, Pnf, idj, gj);
                const uint32_t dw = lds32(rs + F::O_DW + 128u * (uint32_t)jf + lane4);
                const double u = uniform_of(cs0 + jf - cpad);
                double xk[R];
                const double En = eval_state((int)dw, idj, gj, xk);
                const bool accf = (En < Ef) || metropolis_accept(__dsub_rn(Ef, En), u);
                // the chain's own state proposed (1 / S of the state moves): accepted, and nothing changes but the counters
                // (unless E is not the energy of the chain's configuration yet: the initial 1e99, an energy set by the caller)
                if (accf && (int)dw == cstate && __double_as_longlong(En) == __double_as_longlong(Ef) && !(TAB_DBG & 1)) accb |= 1u << jf;
                else if (accf && TAB_REPLAY) {
                    moved = true; st_new = (int)dw; E_mv = En;
#pragma unroll
                    for (int k = 0; k < R; ++k) t_mv[k] = xk[k];
                }
                else if (accf) rare = true;                                                    // accepted after all: the careful path
                if (TAB_STATS) {
                    why |= 4u;
                    atomicAdd(C.spec_stats + 5, 1ull);                                           // resolutions
                    if (rare || moved) atomicAdd(C.spec_stats + 6, 1ull);                        // ... that found an accepted move to another state
                }
            }
            // ---- an accepted move to another state: the other chains of the warp keep their batch; this chain's remaining steps
            //      are replayed by its lane alone, every term evaluated exactly from the new state's rows (its tables are stale).
            //      Anything undecidable in the replay -- a second state move the lower bound cannot reject -- and the whole
            //      batch takes the careful path after all ----
            const uint32_t mvm = TAB_REPLAY ? __ballot_sync(SPEC_FULL, moved) : 0u;
            if (mvm) {
                for (uint32_t m = mvm; m; m &= m - 1) {
                    const int L = __ffs((int)m) - 1;
                    stage_state(L, __shfl_sync(SPEC_FULL, st_new, L), __shfl_sync(SPEC_FULL, lam, L));
                }
                __syncwarp();
                bool fb = false;
                uint32_t ab = 0u, Prs = Psf, Prn = Pnf;
                if (moved) {
                    ab = (accb & ((1u << jf) - 1u)) | (1u << jf);
#pragma unroll 1
                    for (int j = jf + 1; j < BL; ++j) {
                        const uint32_t dwj = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
                        const uint4 w = lds128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                        const double2 txj = lds128(rs + F::O_TX + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                        if (!(E_mv < txj.y) && (int)dwj >= 0 && dwj != TAB_PAD) { fb = true; break; }
                        const int r = (int)((w.z >> 16) & 7u);
                        if (r >= R) continue;                        // (a state move the lower bound rejects, padding)
                        const uint32_t Ps2 = (Prs + w.x) & MS, Pn2 = (Prn + w.y) & MN;
                        int idj[R], gj, ir = 0;
                        true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Ps2, Pn2, idj, gj);
#pragma unroll
                        for (int k = 0; k < R; ++k) ir = r == k ? idj[k] : ir;
                        const double tp = cell_value(lane, r, ir, gj);
                        double xk[R];
#pragma unroll
                        for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t_mv[k];
                        double En = xk[0];
#pragma unroll
                        for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
                        const double Ahi = __dsub_rn(E_mv, txj.x), Alo = __dadd_rn(Ahi, 2.0 * TAB_MARGIN);
                        bool acc = En < Ahi;
                        if (!acc && !(En > Alo)) acc = (En < E_mv) || metropolis_accept(__dsub_rn(E_mv, En), uniform_of(cs0 + j - cpad));
                        if (acc) {
                            E_mv = En;
#pragma unroll
                            for (int k = 0; k < R; ++k) t_mv[k] = xk[k];
                            Prs = Ps2; Prn = Pn2;
                            ab |= 1u << j;
                        }
                    }
                }
                if (__any_sync(SPEC_FULL, fb)) {
                    // (the careful path starts from the batch's first step: the chains' old states back into shared memory)
                    for (uint32_t m = mvm; m; m &= m - 1) {
                        const int L = __ffs((int)m) - 1;
                        stage_state(L, __shfl_sync(SPEC_FULL, cstate, L), __shfl_sync(SPEC_FULL, lam, L));
                    }
                    __syncwarp();
                    rare = true;
                } else if (moved) {
                    E = E_mv; Ps = Prs; Pn = Prn; accb = ab; cstate = st_new; reb = true;
#pragma unroll
                    for (int k = 0; k < R; ++k) t[k] = t_mv[k];
                }
            }
        }
        if (second_look && __any_sync(SPEC_FULL, rare)) {
            // ---- careful batch: step by step from the raw proposal words, every rare case handled exactly ----
            ++n_rollbacks;
            did_careful = true;
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
            trace_t(b, 8);
            // the fast path's step with a vote after each test: whatever it cannot decide is evaluated exactly on the spot, for the
            // lanes concerned (true indices = the tabulator's checkpoint + the positions' progress since)
#pragma unroll 1
            for (int j = 0; j < BL; ++j) {
                const uint32_t dw = lds32(rs + F::O_DW + 128u * (uint32_t)j + lane4);
                const uint4 w = lds128u(rs + F::O_W + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                const double2 txj = lds128(rs + F::O_TX + 512u * (uint32_t)j + 16u * (uint32_t)lane);
                const uint32_t Ps2 = (Ps + w.x) & MS, Pn2 = (Pn + w.y) & MN;
                const uint32_t x = __byte_perm(Ps2, Pn2, w.z);
                const uint32_t e = (x & (0xfu | ((w.z >> 16) & 0x10u))) | ((x >> 4) & 0xf0u);
                double tp = lds64(tabl + w.w + e * 256u);
                const int r = (int)((w.z >> 16) & 7u);
                const uint32_t bs = (((Ps2 + nls) & MS) + mgs) & HS, bn = (((Pn2 + nln) & MN) + mgn) & HN;
                const bool bad = r < R && ((TAB_DBG & 2) || reb || (((w.z >> 20) & 1u) ? ((bs >> (8 * (r & 3))) & 0x80u) != 0u : bn != 0u));
                if (__any_sync(SPEC_FULL, bad)) {             // outside the tabulated window (or stale tables): the exact term
                    if (bad) {
                        int idj[R], gj, ir = 0;
                        true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Ps2, Pn2, idj, gj);
#pragma unroll
                        for (int k = 0; k < R; ++k) ir = r == k ? idj[k] : ir;
                        tp = cell_value(lane, r, ir, gj);
                    }
                }
                double xk[R];
#pragma unroll
                for (int k = 0; k < R; ++k) xk[k] = r == k ? tp : t[k];
                double En = xk[0];
#pragma unroll
                for (int k = 1; k < R; ++k) En = __dadd_rn(En, xk[k]);
                const double Ahi = __dsub_rn(E, txj.x), Alo = __dadd_rn(Ahi, 2.0 * TAB_MARGIN);
                bool acc = En < Ahi;
                const bool unsure = !acc && !(En > Alo);
                if (__any_sync(SPEC_FULL, unsure)) {          // inside the guard band: the specified exact rule
                    if (unsure) acc = (En < E) || metropolis_accept(__dsub_rn(E, En), uniform_of(cs0 + j - cpad));
                }
                // a state move the lower bound cannot reject: evaluated exactly
                const bool lbf = !(E < txj.y) && (int)dw >= 0 && dw != TAB_PAD;
                if (__any_sync(SPEC_FULL, lbf)) {
                    bool accst = false;
                    if (lbf) {
                        int idj[R], gj;
                        true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Ps, Pn, idj, gj);
                        En = eval_state((int)dw, idj, gj, xk);
                        acc = (En < E) || metropolis_accept(__dsub_rn(E, En), uniform_of(cs0 + j - cpad));
                        accst = acc && (int)dw != cstate;     // (the chain's own state proposed: accepted, nothing changes)
                    }
                    // an accepted state move: the chain's record / gamma row in shared memory now, its tables after the batch
                    uint32_t m = __ballot_sync(SPEC_FULL, accst);
                    while (m) {
                        const int L = __ffs((int)m) - 1;
                        m &= m - 1;
                        stage_state(L, __shfl_sync(SPEC_FULL, (int)dw, L), __shfl_sync(SPEC_FULL, lam, L));
                    }
                    if (accst) { reb = true; cstate = (int)dw; }
                    __syncwarp();
                }
                if (acc) {
                    E = En;
#pragma unroll
                    for (int k = 0; k < R; ++k) t[k] = xk[k];
                    Ps = lbf ? Ps : Ps2;
                    Pn = lbf ? Pn : Pn2;
                    accb |= 1u << j;
                }
            }
            trace_t(b, 10);
        }
        if (second_look && __any_sync(SPEC_FULL, reb)) {
            int idx[R], gi;
            true_indices(xw >= 0, xw >= 0 ? xw % 3 : 0, Ps, Pn, idx, gi);
            // rebuild the tables of the chains that changed their state, around their positions at the end of the batch -- once
            // the tabulator has finished batch b - 1 (it is then idle until this batch is published, and nobody else writes tables)
            uint32_t m = __ballot_sync(SPEC_FULL, reb);
            trace_t(b, 9);
            if (m && b >= 1) mbar_wait(b_tab + 8u * (uint32_t)((b - 1) % 3), (uint32_t)((b - 1) / 3) & 1u);
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
                rebat = b;
                // window [P - MRG0, P + MRG0] in every field
                ov_nls = ((((MS - Ps) + 0x01010101u) & MS) + (NL0S & MS)) & MS;
                ov_nln = ((((MN - Pn) + 0x00000101u) & MN) + (NL0 & MN)) & MN;
                ov_mgs = MG0S; ov_mgn = MG0 & 0xffffu;
#pragma unroll
                for (int f = RS; f < 4; ++f) { ov_nls &= ~(0xffu << (8 * f)); ov_mgs |= 127u << (8 * f); }
                if (!HG) { ov_nln = 0u; ov_mgn = 127u | (127u << 8); }
            }
        }
        if (TAB_STATS && second_look && lane == 0) {             // cycles spent on second looks, by outcome (7: resolved only, 8: replay, 9: careful)
            atomicAdd(C.spec_stats + (did_careful ? 9 : (rebw ? 8 : 7)), (unsigned long long)(clock64() - t_sl0));
            atomicAdd(C.spec_stats + (did_careful ? 12 : (rebw ? 11 : 10)), 1ull);
        }
        // ---- publish the batch, release the stage ----------------------------------------------------
        sts32(rs + F::O_PS + lane4, Ps);
        sts32(rs + F::O_PN + lane4, Pn);
        sts32(rs + F::O_ACC + lane4, accb | rebw);
        sts64d(rs + F::O_E + lane8, E);
        trace(b, 3);
        mbar_arrive(b_empty + 8u * (uint32_t)stage);
    }

    if (lane == 0 && n_batches) {
        atomicAdd(C.spec_stats, (unsigned long long)n_batches);
        if (n_rollbacks) atomicAdd(C.spec_stats + 1, (unsigned long long)n_rollbacks);
    }
    if (live) C.E[c] = E;
}

template <int R, bool HG, int NSV>
static cudaError_t launch_tab_ns(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    using F = TabCfg<R, HG, NSV>;
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t smem = (size_t)T.sig4_bytes + (size_t)F::NE * 256 + (HG ? (size_t)32 * T.gh4_stride * 8 : 0) + (size_t)R * 512 + 256 +
                        (size_t)F::NS * F::STAGE + 3 * 6 * 128 + 3 * (size_t)(R + 1) * 128 + 256 + 512 + ((size_t)(R + 1) * 16 * 128 + 128) + 128;
    if (smem + 1024 > (size_t)max_optin) return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(mcmc_tab_kernel<R, HG, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)((C.n_chains + 31) / 32);
    mcmc_tab_kernel<R, HG, NSV><<<grid, 448, smem, st>>>(T, C, A);
    return cudaGetLastError();
}
// a ring of four stages where the tables leave room for it, else three or two (five restraints, long sigma grids)
template <int R, bool HG>
static cudaError_t launch_tab_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    cudaError_t e = launch_tab_ns<R, HG, TAB_NS>(T, C, A, st);
    if (e == cudaErrorNotSupported) e = launch_tab_ns<R, HG, 3>(T, C, A, st);
    if (e == cudaErrorNotSupported) e = launch_tab_ns<R, HG, 2>(T, C, A, st);
    return e;
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
