// K5, warp-specialised speculative Metropolis kernel.
//
// With 4096 chains per GPU a step costs the time ONE warp needs for its dependent instruction stream
// (sampler_spec.cu; DESIGN.md section 4: ~4.5 cycles per shared-memory / shuffle instruction, ~1.5 per other
// instruction).  Only a part of that stream is the chain's critical path -- resolve the pending step, evaluate the
// next proposal under both outcomes, commit.  Everything else is moved to a second warp of the same CTA (same SM
// sub-partition: it issues in the first warp's stalls):
//
//   consumer warp (one per 32 / L chains, L lanes per chain, one lane per restraint)
//       resolve -> evaluate -> commit, with the speculation, the exact reciprocal-table division and the rollback of
//       a batch to a careful step-by-step path of sampler_spec.cu.  It reads PRE-DECODED per-lane step words from a
//       shared-memory ring -- the shared-memory addresses of the record / gamma value to read, the sigma-entry and
//       gamma offsets already scaled to bytes -- decides in fp64 against thresholds the producer prepared
//       (E - (ln u +- margin): one DSETP after the energy sum), keeps its indices scaled (24 sigma_index,
//       8 gamma_index) so that an address is one add, and publishes per step its packed indices and per batch the
//       accept bits and the energy.
//   generator warp (one per consumer warp)
//       generates the Philox blocks of the next batch (one block per lane and L steps), decodes them into the
//       per-lane words, requests the proposed states' records / gamma windows with cp.async into the ring's slots.
//   accounting warp (one per consumer warp)
//       turns the batches the consumer has finished into the bookkeeping: run-length histograms from the
//       published indices (compacted per batch, one RED per event), acceptance counters from {proposal word, accept
//       bit}, state trace, and the snapshots that fall on the last step of a batch (the kernel front-pads the first
//       batch so that they do whenever traj_every is a multiple of BL) -- none of this is on the consumer's stream.
//
// Hand-off: per consumer warp a ring of NS stages (one stage = one batch of BL steps), three mbarriers per stage.
// full[stage]: 32 generator lanes arrive after their stores + 32 cp.async-completion arrivals (cp.async.mbarrier.
// arrive.noinc), so a completed phase means the ring words AND the prefetched slots have landed.  empty[stage]:
// the 32 consumer lanes arrive when the batch is done.  free[stage]: the accountant has copied the batch's words
// into registers (it then works from the copy while the generator refills the stage).  The consumer needs the stage after the current one only
// for the LAST step of a batch (it evaluates the first step of the next batch) and waits for it there, so two
// stages are enough for the two warps to overlap.
// Same RNG stream, same fp64 operation order, same outputs as every other Metropolis kernel (bit for bit).
#include "spec_common.cuh"

namespace bcp {

#ifndef WS_NS
#define WS_NS 4                    // ring stages
#endif
#ifndef WS_BLV
#define WS_BLV 4                   // steps per batch (a multiple of L, at most 8: the accept bits of a batch are one byte)
#endif
#define WS_BL(L) ((L) > WS_BLV ? (L) : WS_BLV)
#ifndef WS_ROLLED
#define WS_ROLLED 0                // 1: the steps of a fast batch as a rolled loop (every step its own basic block)
#endif
#ifndef WS_TIMING_NO_REPLAY
#define WS_TIMING_NO_REPLAY 0      // 1: timing experiments only (scripts/tune_ws.sh) -- the producer skips the bookkeeping, results are WRONG
#endif

template <int L>
struct WsCfg {
    static constexpr int BL = WS_BL(L);          // steps per batch = per ring stage
    static constexpr int NB = BL / L;            // Philox blocks a producer lane generates per batch
    static constexpr int NS = WS_NS;
    static constexpr int CPW = 32 / L;           // chains per consumer warp
    static constexpr int NSLOT = NS * BL;        // proposal slots per chain
    static_assert(BL % L == 0 && BL <= 8 && NS >= 2, "batch / ring shape");
    // byte offsets inside a ring stage (per consumer warp)
    static constexpr int O_LW = 0;                           // [BL][32] per-LANE step words {w, recaddr, nbase, ln u}
    static constexpr int O_TH = BL * 512;                    // [BL][CPW] {ln u + margin, ln u - margin} as doubles (fast accept test)
    static constexpr int O_U = O_TH + BL * CPW * 16;         // [BL][CPW] Metropolis uniform (careful path)
    static constexpr int O_E = O_U + BL * CPW * 8;           // [CPW] energy after the batch                 (consumer -> producer)
    static constexpr int O_IDX = O_E + CPW * 8;              // [BL][32] 24 sigma_index | 8 gamma_index << 16 after each step (consumer -> producer)
    static constexpr int O_DW = O_IDX + BL * 128;            // [BL][CPW] raw proposal word (bookkeeping, careful path)
    static constexpr int O_IST = O_DW + BL * CPW * 4;        // [BL][CPW] proposed state
    static constexpr int O_WB = O_IST + BL * CPW * 4;        // [CPW] 8 x first gamma-row entry of the stage's prefetched windows
    static constexpr int O_ACC = O_WB + CPW * 4;             // [CPW] accept bits of the batch               (consumer -> producer)
    static constexpr int STAGE = O_ACC + CPW * 4;
    static constexpr int EVQ = BL * 32 * 8;                  // per producer warp: histogram events of one batch
};
// per-lane step word w: bit 31 nuisance move, bits 15:8 = 8 (dgamma + 1) in the gamma lane (8 elsewhere),
// bits 7:0 = 24 (dsigma + 1) of the lane's restraint

__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint32_t a) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WS_DONE_%=;\n"
        "bra WS_WAIT_%=;\n"
        "WS_DONE_%=:\n"
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

template <int L, int R, bool HG, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
mcmc_ws_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
               const __grid_constant__ LaunchArgs A) {
    using W = WsCfg<L>;
    constexpr int BL = W::BL;
    constexpr int NB = W::NB;
    constexpr int NS = W::NS;
    constexpr int CPW = W::CPW;
    constexpr int EB = SPEC_ENT_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(8) uint64_t bars[8][3][NS];         // [consumer warp][0 full, 1 empty, 2 free][stage]

    // Warp roles (0 consumer, 1 generator, 2 accountant, 3 idle).  A warp runs on SM sub-partition warp % 4, and a consumer
    // that has its sub-partition to itself is 25 % faster than one that shares it with its two helpers (measured: 704
    // chains 3.30 ms against 4.28 ms for the spec kernel, 4096 chains 4.72 against 4.59), so:
    //   1 consumer warp per CTA :  C G A                      (three sub-partitions)
    //   2 consumer warps        :  C0 C1 G0 A0 -  -  G1 A1    (256 threads, two idle warps: the helpers share 2 and 3)
    //   3, 4 consumer warps     :  C0.. G0.. A0..             (every sub-partition hosts a consumer and its helpers)
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int NW = blockDim.x == 256 ? 2 : (int)(blockDim.x / 96);
    int role, cw;
    if (NW == 2) {
        role = warp < 2 ? 0 : ((warp == 4 || warp == 5) ? 3 : 1 + (warp & 1));
        cw = warp < 2 ? warp : (warp >> 2) & 1;
    } else {
        role = warp / NW;
        cw = warp - role * NW;
    }
    if ((int)threadIdx.x < NW) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            mbar_init(smem_u32(&bars[threadIdx.x][0][k]), 64);   // generator: 32 arrivals + 32 cp.async completions
            mbar_init(smem_u32(&bars[threadIdx.x][1][k]), 32);   // consumer: batch done
            mbar_init(smem_u32(&bars[threadIdx.x][2][k]), 32);   // accountant: the stage's words are copied out, it may be refilled
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_sigma_table(reinterpret_cast<double*>(smem_raw), T.sig4, T.sig4_bytes, &tma_bar);     // has the block barrier
    if (role == 3) return;

    const long long n = C.n_chains;
    const int q = lane & (L - 1);                            // lane within the chain's group
    const int jg = lane / L;                                 // chain within the warp
    long long c = ((long long)blockIdx.x * NW + cw) * CPW + jg;
    const bool live = c < n;                                 // dead groups shadow the last chain, no stores
    if (!live) c = n - 1;
    const int qr = q < R ? q : 0;                            // restraint this lane evaluates / follows (idle lanes: 0)
    const bool owner = live && q < R;
    const bool lead = live && q == 0;
    const bool glane = HG && q == R - 1;                     // lane of the gamma restraint
    const uint32_t own_mask = q < R ? (1u << (20 + q)) : 0u;
    const int fsh = 2 * qr;
    constexpr int GL = R < L ? R : 0;                        // lane that keeps the gamma histogram (an idle lane if any)
    const bool ghist = HG && live && q == GL;

    const int lam = C.lam[c];
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double* __restrict__ rec = T.rec;
    const double* __restrict__ gh4 = HG ? T.gsse_halo4 : nullptr;
    const int G = HG ? T.r[R - 1].n_gamma : 1;
    const int GS = HG ? T.gh4_stride : 0;
    const int nsig = T.r[qr].n_sigma;
    const int s_end = (int)(A.s_end - A.s_begin);
    const long long burn_l = A.burn - A.s_begin;
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }
    // front padding: batch b covers the local steps b BL - pad .. b BL - pad + BL - 1, so that the first snapshot
    // (and, with traj_every a multiple of BL, every snapshot) is the LAST step of its batch; padded steps
    // (s < 0 or s >= s_end) carry ln u = +inf and are never accepted
    const int pad = next_snap >= 0 ? BL - 1 - next_snap % BL : 0;
    const int nbatch = (s_end + pad + BL - 1) / BL;

    const SpecLayout ly = spec_layout(W::NSLOT, R, HG, GS);
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t chain_base = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(cw * CPW + jg) * (uint32_t)ly.chain_bytes;
    const uint32_t cur_rec = chain_base + ly.cur_rec + 16 * qr;
    const uint32_t cur_row = chain_base + ly.cur_row;
    const uint32_t slots = chain_base + ly.slot0;
    const uint32_t ring = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(NW * CPW) * (uint32_t)ly.chain_bytes +
                          (uint32_t)(cw * NS) * (uint32_t)W::STAGE;
    const uint32_t bar_full = smem_u32(&bars[cw][0][0]), bar_empty = smem_u32(&bars[cw][1][0]), bar_free = smem_u32(&bars[cw][2][0]);
    auto slot_of = [&](int stage, int j) { return slots + (uint32_t)(stage * BL + j) * (uint32_t)ly.slot_bytes; };
    auto ring_of = [&](int stage) { return ring + (uint32_t)stage * (uint32_t)W::STAGE; };
    const uint32_t idx_glane = W::O_IDX + 4u * (uint32_t)(jg * L + (HG ? R - 1 : 0));   // the chain's gamma lane in an IDX row

    if (role == 1) {
        // =====================================================================================
        // generator: Philox + decode + prefetch, one batch ahead of the consumer
        // =====================================================================================
        const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
        const uint32_t S = (uint32_t)T.S;
        const uint32_t thresh = T.nuis_thresh;
        const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
        const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
        const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);

        uint32_t gm8 = HG ? (uint32_t)(8 * C.igm[(size_t)(R - 1) * n + c]) : 0u;   // the chain's scaled gamma index (every lane)

        uint32_t rk0[10], rk1[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            rk0[i] = seed_lo + (uint32_t)i * BCP_PHILOX_W0;
            rk1[i] = seed_hi + (uint32_t)i * BCP_PHILOX_W1;
            asm volatile("" : "+r"(rk0[i]), "+r"(rk1[i]));
        }
        // per-lane constants of the step words: w = nuis | ((glane ? 8 f_gamma : 8) << 8) | 24 f_sigma, f = delta + 1
        const uint32_t wg_mul = glane ? 2048u : 0u, wg_add = glane ? 0u : 2048u;
        const uint32_t nb_row = cur_row + 24u;               // + 8 f_gamma: entry of gamma index + (delta) in the halo-4 row

        // one batch: lane (chain jg, position q + m L) generates that step; then every lane writes ITS words of
        // every step of the batch
        auto generate = [&](int b, int stage) {
            const uint32_t rs = ring_of(stage);
            uint32_t g_dw[NB];
            float g_lu[NB];
            const uint32_t wbase = (gm8 >> 3) & ~1u;         // first halo-4 row entry of this stage's prefetched windows
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const int t = m * L + q;
                const int s = b * BL - pad + t;
                const unsigned long long ctr = ctr0 + (unsigned long long)(long long)s;
                uint32_t pc0 = (uint32_t)ctr, pc1 = (uint32_t)(ctr >> 32), pc2 = ch_lo, pc3 = ch_hi;
#pragma unroll
                for (int i = 0; i < 10; ++i) {
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
                const uint32_t a = (uint32_t)(z1 >> 32);
                const uint32_t bb = (HG && hi == (uint32_t)(R - 1)) ? (uint32_t)(z2 >> 32) : 1u;
                const uint32_t dw = nuis ? ((0x80000000u | SPEC_DW_IDLE | (0x00100000u << hi)) ^ ((a ^ 1u) << (2 * hi)) ^ ((bb ^ 1u) << 16))
                                         : SPEC_DW_IDLE;
                const double u = spec_u2(pc2, pc3);
                g_dw[m] = dw;
                g_lu[m] = (s >= 0 && s < s_end) ? __logf((float)u) : __int_as_float(0x7f800000);
                const uint32_t e = (uint32_t)(t * CPW + jg);
                sts32(rs + W::O_DW + 4u * e, dw);
                {
                    const double lud = (double)g_lu[m];       // +inf for padding: both thresholds +inf, the step is a sure reject
                    sts128(rs + W::O_TH + 16u * e, make_double2(__dadd_rn(lud, (double)SPEC_MARGIN), __dsub_rn(lud, (double)SPEC_MARGIN)));
                }
                sts64d(rs + W::O_U + 8u * e, u);
                sts32(rs + W::O_IST + 4u * e, hi);
                if (!nuis) {
                    const uint32_t slot = slot_of(stage, t);
                    const double* r2 = rec + (size_t)hi * (2 * R);
#pragma unroll
                    for (int k = 0; k < R; ++k) cp_async16(slot + 16 * k, r2 + 2 * k);
                    cp_async8(slot + ly.slot_c, energy + hi);
                    if (HG) {
                        const double* w = gh4 + (size_t)hi * GS + wbase;
#pragma unroll
                        for (int k = 0; k < 5; ++k) cp_async16(slot + ly.slot_win + 16 * k, w + 2 * k);
                    }
                }
            }
            const uint32_t slot_q = slot_of(stage, 0) + 16u * (uint32_t)qr;
            // window entry of halo-4 row entry k: k - wbase  ->  value of gamma index gm: slot_win + 8 (gm + 4 - wbase)
            const uint32_t nb_win = slot_of(stage, 0) + (uint32_t)ly.slot_win + 32u - 8u * wbase;
#pragma unroll
            for (int t = 0; t < BL; ++t) {
                const uint32_t dwt = __shfl_sync(SPEC_FULL, g_dw[t / L], t % L, L);
                const float lut = __shfl_sync(SPEC_FULL, g_lu[t / L], t % L, L);
                const bool nuis = (int)dwt < 0;
                const uint32_t fs = (dwt >> fsh) & 3u, fg = (dwt >> 16) & 3u;
                const uint32_t so = (uint32_t)t * (uint32_t)ly.slot_bytes;
                const uint32_t recaddr = nuis ? cur_rec : slot_q + so;
                uint32_t nbase = recaddr;
                if (glane) nbase = nuis ? nb_row + 8u * fg : nb_win + so;
                const uint32_t w = (dwt & 0x80000000u) | (fs * (uint32_t)EB + fg * wg_mul + wg_add);
                sts128u(rs + W::O_LW + 512u * (uint32_t)t + 16u * (uint32_t)lane, w, recaddr, nbase, __float_as_uint(lut));
            }
            if (q == 0) sts32(rs + W::O_WB + 4u * (uint32_t)jg, 8u * wbase);
            mbar_arrive_cp_async(bar_full + 8u * (uint32_t)stage);
            mbar_arrive(bar_full + 8u * (uint32_t)stage);
        };

        // batch i goes into stage i mod NS once the accountant has copied the words of batch i - NS out of it (which implies
        // that the consumer is done with it).  Batch nbatch is generated too: the consumer's last step evaluates its first
        // proposal (padding).  The gamma window of the prefetches is placed around the chain's gamma index at the end of
        // the stage's previous batch (the consumer's published indices are still there).
        int stage = 0, round = 0;                             // i == round * NS + stage
        for (int i = 0; i <= nbatch; ++i) {
            if (i >= NS) {
                mbar_wait(bar_free + 8u * (uint32_t)stage, (uint32_t)(round - 1) & 1u);
                if (HG) gm8 = lds32(ring_of(stage) + idx_glane + 128u * (uint32_t)(BL - 1)) >> 16;
            }
            generate(i, stage);
            if (++stage == NS) { stage = 0; ++round; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }
    if (role == 2) {
        // =====================================================================================
        // accountant: the bookkeeping of the batches the consumer has finished
        // =====================================================================================
        const int grp = C.group[c];
        unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
        const uint32_t sbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)T.r[qr].hist_sigma;
        const uint32_t gbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)(HG ? T.r[R - 1].hist_gamma : 0);
        const bool tracing = A.state_trace != nullptr;
        const bool traced = tracing && lead && c < A.trace_n;
        const uint32_t evq = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(NW * CPW) * (uint32_t)ly.chain_bytes +
                             (uint32_t)(NW * NS) * (uint32_t)W::STAGE + (uint32_t)cw * (uint32_t)W::EVQ;
        const uint32_t lane_lt = (1u << lane) - 1u;

        int state = C.state[c];
        uint32_t sg24 = (uint32_t)(EB * C.isg[(size_t)qr * n + c]);     // this lane's scaled sigma index as last published
        uint32_t gm8 = HG ? (uint32_t)(8 * C.igm[(size_t)(R - 1) * n + c]) : 0u;   // the chain's scaled gamma index (every lane)
        int since_sg = burn, since_gm = burn, since_state = burn;
        unsigned int acc_own = 0, acc_state = 0, acc_total = 0;

        // accepted state move of a finished step (rare on large ensembles)
        auto account_state = [&](int s, int i2) {
            ++acc_state;
            if (i2 != state) {
                if (s > burn && lead) atomicAdd(g_state_counts + state, (unsigned long long)(s - since_state));
                if (s > burn) since_state = s;
                state = i2;
            }
        };
        auto next_snapshot = [&]() {
            ++snap_idx;
            const long long ns = (long long)next_snap + A.traj_every;
            next_snap = ns < s_end ? (int)ns : -1;
        };

        // the bookkeeping of a finished batch: run lengths of the indices the consumer published after every step,
        // counters from {proposal word, accept bit}; the batch's histogram events are compacted into a per-warp queue
        // and flushed with one RED each
        auto account = [&](int b, int stage) {
            const uint32_t rs = ring_of(stage);
            // copy the batch's words out of the stage, then hand the stage back to the generator
            uint32_t w_own[BL], w_gam[BL], w_dw[BL], w_ist[BL];
            const uint32_t accb = lds32(rs + W::O_ACC + 4u * (uint32_t)jg);
            const double E_end = lds64(rs + W::O_E + 8u * (uint32_t)jg);
#pragma unroll
            for (int t = 0; t < BL; ++t) {
                w_own[t] = lds32(rs + W::O_IDX + 128u * (uint32_t)t + 4u * (uint32_t)lane);
                w_gam[t] = HG ? lds32(rs + idx_glane + 128u * (uint32_t)t) : 0u;
                w_dw[t] = lds32(rs + W::O_DW + 4u * (uint32_t)(t * CPW + jg));
                w_ist[t] = lds32(rs + W::O_IST + 4u * (uint32_t)(t * CPW + jg));
            }
            mbar_arrive(bar_free + 8u * (uint32_t)stage);
#ifdef WS_TIMING_ACC_COPY_ONLY
            {   // timing experiment: keep the loads alive, skip the bookkeeping
                uint32_t x = accb ^ (uint32_t)__double2loint(E_end);
#pragma unroll
                for (int t = 0; t < BL; ++t) x ^= w_own[t] ^ w_gam[t] ^ w_dw[t] ^ w_ist[t];
                if (x == 0x12345678u) acc_total++;
                return;
            }
#endif
            const int sb = b * BL - pad;
            uint32_t evn = 0;
#pragma unroll
            for (int t = 0; t < BL; ++t) {
                const int s = sb + t;
                const uint32_t dw = w_dw[t];
                const uint32_t mine = w_own[t] & 0xffffu;
                const bool acc = (accb >> t) & 1u;
                const bool rec_on = s > burn;
                acc_total += acc ? 1u : 0u;
                acc_own += (acc && (dw & own_mask)) ? 1u : 0u;
                const bool chs = rec_on && mine != sg24;
                uint32_t bin = sbin0 + sg24 / (uint32_t)EB;
                int since = since_sg;
                bool ev = chs && owner;
                since_sg = chs ? s : since_sg;
                sg24 = mine;
                if (HG) {
                    const uint32_t gnow = w_gam[t] >> 16;
                    const bool chg = rec_on && gnow != gm8;
                    const bool evg = chg && ghist;
                    bin = evg ? gbin0 + (gm8 >> 3) : bin;
                    since = evg ? since_gm : since;
                    ev = ev || evg;
                    since_gm = chg ? s : since_gm;
                    gm8 = gnow;
                }
                const uint32_t cnt = ev ? (uint32_t)(s - since) : 0u;
                const uint32_t m = __ballot_sync(SPEC_FULL, cnt != 0u);
                if (cnt != 0u) sts64u(evq + 8u * (evn + (uint32_t)__popc(m & lane_lt)), bin, cnt);
                evn += (uint32_t)__popc(m);
                if (acc && (int)dw >= 0) account_state(s, (int)w_ist[t]);
                if (tracing) {
                    if (traced && s >= burn && s < s_end) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
                }
            }
            __syncwarp();
            for (uint32_t k = (uint32_t)lane; k < evn; k += 32) {
                const uint2 e = lds64u(evq + 8u * k);
                atomicAdd(C.param_counts + e.x, (unsigned long long)e.y);
            }
            __syncwarp();
            // snapshots of this batch: the one on the last step is stored here (indices and state == the batch's final
            // ones), the others were stored by the consumer's careful path
            while (next_snap >= 0 && next_snap < sb + BL) {
                if (next_snap == sb + BL - 1) {
                    const size_t o = (size_t)snap_idx * n + c;
                    if (lead) {
                        A.traj_E[o] = E_end;
                        A.traj_state[o] = state;
                        A.traj_accept[o] = (accb >> (BL - 1)) & 1u;
                    }
                    if (owner) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = (int)(sg24 / (uint32_t)EB);
                    if (HG && glane && live) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = (int)(gm8 >> 3);
                }
                next_snapshot();
            }
        };

        int stage = 0, round = 0;                             // b == round * NS + stage
        for (int b = 0; b < nbatch; ++b) {
            mbar_wait(bar_empty + 8u * (uint32_t)stage, (uint32_t)round & 1u);
#if WS_TIMING_NO_REPLAY
            mbar_arrive(bar_free + 8u * (uint32_t)stage);
#else
            account(b, stage);
#endif
            if (++stage == NS) { stage = 0; ++round; }
        }
        // ---- flush run-length counters, persist the chain (E is persisted by the consumer) ----
        if (lead) {
            if (s_end > since_state) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - since_state));
            C.state[c] = state;
            C.accepted[c] += acc_total;
            C.sep[(size_t)R * n + c] += acc_state;
        }
        if (owner) {
            if (s_end > since_sg) atomicAdd(C.param_counts + sbin0 + sg24 / (uint32_t)EB, (unsigned long long)(s_end - since_sg));
            C.isg[(size_t)q * n + c] = (int)(sg24 / (uint32_t)EB);
            C.sep[(size_t)q * n + c] += acc_own;
        }
        if (ghist) {
            if (s_end > since_gm) atomicAdd(C.param_counts + gbin0 + (gm8 >> 3), (unsigned long long)(s_end - since_gm));
            C.igm[(size_t)(R - 1) * n + c] = (int)(gm8 >> 3);
        }
        return;
    }

    // =========================================================================================
    // consumer: the chain's critical path
    // =========================================================================================
    const double logZ = T.logZ[lam];
    const uint32_t ent0 = smem0 + (uint32_t)EB * (uint32_t)(T.ent_off[qr] + 2);      // entry of sigma index 0
    const uint32_t entm = ent0 - (uint32_t)EB;                                       // ... minus the bias of the step word's field
    const double cnorm = T.r[qr].cnorm;
    const int nsig24 = EB * nsig, G8 = 8 * G;
    const uint32_t lw_lane = W::O_LW + 16u * (uint32_t)lane;
    const uint32_t idx_lane = W::O_IDX + 4u * (uint32_t)lane;
    auto bcast_d = [&](double v, int src) { return __shfl_sync(SPEC_FULL, v, src, L); };

    int state = C.state[c];
    double E = C.E[c];
    int sg24 = EB * C.isg[(size_t)qr * n + c];               // 24 x this lane's sigma index
    int gm8 = glane ? 8 * C.igm[(size_t)(R - 1) * n + c] : 0;  // 8 x gamma index, in the gamma lane only (0 elsewhere)
    // lane 0 adds energy + logZ to its term (reference order (C + t_0) + t_1 ...); x + (-0.0) == x bitwise
    double Cst = q == 0 ? __dadd_rn(energy[state], logZ) : -0.0;
    unsigned int n_batches = 0, n_rollbacks = 0;

    if (q < R) sts128(cur_rec, *reinterpret_cast<const double2*>(rec + ((size_t)state * R + q) * 2));
    if (HG) {
        const double2* src = reinterpret_cast<const double2*>(gh4 + (size_t)state * GS);
        for (int k = q; k < GS / 2; k += L) sts128(cur_row + 16 * k, src[k]);
    }
    __syncwarp();

    // ---- this lane's contribution to the energy of a proposal (per-lane step words lw): its restraint term, plus
    //      energy + logZ in lane 0.  A: on the committed state, B: on the committed state shifted by the pending
    //      step (a0 = 24 dsigma, g0x = 8 dgamma).  Window misses are only flagged (the careful path re-evaluates).
    //      wb8 = 8 x first row entry of the proposal's prefetched gamma window (24 outside the gamma lane: never a miss).
    struct Cand { double xA, xB; int a1, g1x; bool miss; };
    auto evaluate = [&](const uint4& lw, int wb8, int a0_, int g0x_) {
        Cand k;
        const bool nuis_n = (int)lw.x < 0;
        const uint32_t ab = lw.x & 0xffu;
        k.a1 = (int)ab - EB;
        k.g1x = (int)((lw.x >> 8) & 0xffu) - 8;
        const uint32_t eA = entm + (uint32_t)sg24 + ab;      // unwrapped: entries carry a halo of 2
        const uint32_t eB = eA + (uint32_t)a0_;
        const double nlA = lds64(eA), dvA = lds64(eA + 8), yA = lds64(eA + 16);
        const double nlB = lds64(eB), dvB = lds64(eB + 8), yB = lds64(eB + 16);
        const uint32_t recaddr = lw.y;
        const double2 rv = lds128(recaddr);
        double numA = rv.x, numB = rv.x;
        k.miss = false;
        if (HG) {
            // the other lanes re-read their own sse word: nbase == recaddr, gm8 == g0x == 0 there.  A state-move proposal
            // reads the window entry of the current gamma index; its neighbours (+- 1) must be inside the 10-wide window
            k.miss = !nuis_n && (uint32_t)(gm8 + 24 - wb8) > 56u;
            uint32_t na = lw.z + (uint32_t)gm8;
            na = k.miss ? recaddr : na;
            numA = lds64(na);
            numB = lds64(na + (uint32_t)g0x_);
        }
        double Cn = Cst;
        if (!nuis_n && q == 0) Cn = __dadd_rn(lds64(recaddr + 16 * R), logZ);
        double t = __dadd_rn(nlA, div_markstein(numA, dvA, yA));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, rv.y);
        k.xA = __dadd_rn(t, Cn);
        t = __dadd_rn(nlB, div_markstein(numB, dvB, yB));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, rv.y);
        k.xB = __dadd_rn(t, Cn);
        return k;
    };
    // the same on the committed state only, from the raw proposal word; the gamma value of a proposed state (i2)
    // is read from global memory
    auto evaluate_exact = [&](uint32_t dn, uint32_t slot, uint32_t i2) {
        const bool nuis_n = (int)dn < 0;
        const int d1 = (int)((dn >> fsh) & 3u) - 1, g1 = (int)((dn >> 16) & 3u) - 1;
        const uint32_t eA = ent0 + (uint32_t)(sg24 + EB * d1);
        const double nlA = lds64(eA), dvA = lds64(eA + 8), yA = lds64(eA + 16);
        const double2 rv = lds128(nuis_n ? cur_rec : slot + 16 * qr);
        double num = rv.x;
        if (HG) {
            if (glane) num = nuis_n ? lds64(cur_row + (uint32_t)(gm8 + 8 * (4 + g1))) : __ldg(gh4 + (size_t)i2 * GS + ((gm8 >> 3) + 4));
        }
        double Cn = Cst;
        if (!nuis_n && q == 0) Cn = __dadd_rn(lds64(slot + ly.slot_c), logZ);
        double t = __dadd_rn(nlA, div_markstein(num, dvA, yA));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, rv.y);
        return __dadd_rn(t, Cn);
    };
    // commit of a decision on the pending step (a0, g0x): energy and scaled indices only -- the bookkeeping is the
    // producer's
    int a0, g0x;
    auto commit = [&](bool acc, double En) {
        E = acc ? En : E;
        int v = sg24 + a0;                                    // a0 == 0 unless this lane's restraint moves
        v += v < 0 ? nsig24 : 0;
        v -= v >= nsig24 ? nsig24 : 0;
        sg24 = acc ? v : sg24;
        if (HG) {
            int w = gm8 + g0x;                                // g0x == 0 outside the gamma lane
            w += w < 0 ? G8 : 0;
            w -= w >= G8 ? G8 : 0;
            gm8 = acc ? w : gm8;
        }
    };
    auto pending_from_raw = [&](uint32_t dn) {
        a0 = EB * ((int)((dn >> fsh) & 3u) - 1);
        g0x = glane ? 8 * ((int)((dn >> 16) & 3u) - 1) : 0;
    };
    auto publish_step = [&](uint32_t rs, int j) {             // this lane's indices after step j (the producer's bookkeeping)
        sts32(rs + idx_lane + 128u * (uint32_t)j, (uint32_t)sg24 | ((uint32_t)gm8 << 16));
    };

    // ---- prologue: the first pending step ----------------------------------------------------------
    bool nuis_s;
    float lu_s;
    double x_prop;
    mbar_wait(bar_full, 0u);
    {
        const uint32_t dw0 = lds32(ring_of(0) + W::O_DW + 4u * (uint32_t)jg);
        nuis_s = (int)dw0 < 0;
        lu_s = __uint_as_float(lds32(ring_of(0) + lw_lane + 12u));
        pending_from_raw(dw0);
        x_prop = evaluate_exact(dw0, slot_of(0, 0), lds32(ring_of(0) + W::O_IST + 4u * (uint32_t)jg));
    }
    auto wb_of = [&](uint32_t rs) { return glane ? (int)lds32(rs + W::O_WB + 4u * (uint32_t)jg) : 24; };
    int wb_cur = wb_of(ring_of(0));

    int stage = 0;
    uint32_t par = 0;                                         // parity of the current stage's full phase
    for (int b = 0; b < nbatch; ++b) {
        const int s0 = b * BL - pad;
        const int stage_n = stage + 1 == NS ? 0 : stage + 1;
        const uint32_t par_n = stage_n == 0 ? par ^ 1u : par;
        const uint32_t rs = ring_of(stage), rn = ring_of(stage_n);
        // the last step of a batch evaluates the next batch's first step: the next stage is waited for only there
        bool have_next = false;
        int wb_n = 24;
        auto need_next = [&]() {
            if (!have_next) {
                mbar_wait(bar_full + 8u * (uint32_t)stage_n, par_n);
                wb_n = wb_of(rn);
                have_next = true;
            }
        };
        uint32_t accb = 0;

        // ---- fast batch: BL steps on registers only (shared memory is read; written only to publish the indices);
        //      anything rare only raises a flag and the batch is then redone carefully ------------------------
        bool careful = (unsigned)(next_snap - s0) < (unsigned)(BL - 1);   // a snapshot before the batch's last step
        if (!careful) {
            const double sv_E = E, sv_x = x_prop;
            const int sv_sg = sg24, sv_gm = gm8, sv_a0 = a0, sv_g0 = g0x;
            const bool sv_nuis = nuis_s;
            const float sv_lu = lu_s;
            bool rare = false;
#if WS_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int j = 0; j < BL; ++j) {
                const bool last = j + 1 == BL;
                if (last) need_next();
                // the step words of the proposal evaluated in this step (position j + 1, or the next batch's first)
                const uint4 lwj = lds128u((last ? rn : rs + 512u * (uint32_t)(j + 1)) + lw_lane);
                // thresholds of the pending step (position j of this stage) against the committed energy
                const double2 th = lds128(rs + W::O_TH + 16u * (uint32_t)jg + (uint32_t)(16 * CPW) * (uint32_t)j);
                // (1) resolve the pending step: energy in the reference's summation order, accept against E - (ln u +- margin)
                double xv[R];
#pragma unroll
                for (int r = 0; r < R; ++r) xv[r] = bcast_d(x_prop, r);
                const double Ahi = __dsub_rn(E, th.x), Alo = __dsub_rn(E, th.y);
                double En = xv[0];
#pragma unroll
                for (int r = 1; r < R; ++r) En = __dadd_rn(En, xv[r]);
                const bool acc = En < Ahi;
                const bool unsure = !acc && !(En > Alo);                 // guard band, also NaN
                // (2) contributions to the next step under both outcomes of the pending one
                const Cand k = evaluate(lwj, last ? wb_n : wb_cur, a0, g0x);
                rare = rare || unsure || (acc && !nuis_s) || k.miss;
                // (3) commit, publish the indices
                commit(acc, En);
                publish_step(rs, j);
                accb |= acc ? (1u << j) : 0u;
                x_prop = acc ? k.xB : k.xA;
                nuis_s = (int)lwj.x < 0; lu_s = __uint_as_float(lwj.w); a0 = k.a1; g0x = k.g1x;
            }
            if (__any_sync(SPEC_FULL, rare)) {               // roll the batch back
                E = sv_E; x_prop = sv_x; sg24 = sv_sg; gm8 = sv_gm; a0 = sv_a0; g0x = sv_g0;
                nuis_s = sv_nuis; lu_s = sv_lu;
                accb = 0;
                careful = true;
                ++n_rollbacks;
            }
            ++n_batches;
        }

        // ---- careful batch: step by step from the raw proposal words, every rare case handled exactly -------
        if (careful) {
#pragma unroll 1
            for (int j = 0; j < BL; ++j) {
                const int s = s0 + j;
                const bool nb = j + 1 == BL;
                if (nb) need_next();
                const int jn = nb ? 0 : j + 1;
                const uint32_t en = (uint32_t)(CPW * jn + jg), ej = (uint32_t)(CPW * j + jg);
                const uint32_t rsn = nb ? rn : rs;
                const uint32_t dw_n = lds32(rsn + W::O_DW + 4u * en);
                const float lu_n = __uint_as_float(lds32(rsn + 512u * (uint32_t)jn + lw_lane + 12u));
                const uint32_t i2n = lds32(rsn + W::O_IST + 4u * en);
                const uint32_t i2 = lds32(rs + W::O_IST + 4u * ej);
                double En = bcast_d(x_prop, 0);
#pragma unroll
                for (int r = 1; r < R; ++r) En = __dadd_rn(En, bcast_d(x_prop, r));
                const double d = __dsub_rn(E, En);
                const float tt = (float)d - lu_s;
                const bool unsure = !(tt > SPEC_MARGIN) && !(tt < -SPEC_MARGIN);
                bool acc = tt > SPEC_MARGIN;
                if (__any_sync(SPEC_FULL, unsure)) {                      // the specified exact rule, inside the guard band only
                    const double u = lds64(rs + W::O_U + 8u * ej);
                    const bool exact = s >= 0 && s < s_end && ((En < E) || (d >= BCP_ACCEPT_CUTOFF && u < exp_det(d)));
                    acc = unsure ? exact : acc;
                }
                commit(acc, En);
                publish_step(rs, j);
                accb |= acc ? (1u << j) : 0u;
                const bool accst = acc && !nuis_s;
                if (__any_sync(SPEC_FULL, accst)) {
                    if (accst) {       // adopt the slot as the current state, reload the gamma row
                        const uint32_t slot = slot_of(stage, j);
                        if (q < R) sts128(cur_rec, lds128(slot + 16 * qr));
                        if (q == 0) Cst = __dadd_rn(lds64(slot + ly.slot_c), logZ);
                        state = (int)i2;
                        if (HG) {
                            const double2* src = reinterpret_cast<const double2*>(gh4 + (size_t)i2 * GS);
                            for (int kk = q; kk < GS / 2; kk += L) cp_async16(cur_row + 16 * kk, src + kk);
                        }
                    }
                    if (HG) {
                        cp_async_commit();
                        cp_async_wait<0>();
                    }
                    __syncwarp();
                }
                x_prop = evaluate_exact(dw_n, slot_of(nb ? stage_n : stage, jn), i2n);
                if (s >= burn && s < s_end && s == next_snap && j < BL - 1) {       // (the last position is the producer's)
                    const size_t o = (size_t)snap_idx * n + c;
                    if (lead) {
                        A.traj_E[o] = E;
                        A.traj_state[o] = state;
                        A.traj_accept[o] = acc ? 1 : 0;
                    }
                    if (owner) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = sg24 / EB;
                    if (HG && glane && live) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = gm8 >> 3;
                    ++snap_idx;
                    const long long ns = (long long)next_snap + A.traj_every;
                    next_snap = ns < s_end ? (int)ns : -1;
                }
                nuis_s = (int)dw_n < 0; lu_s = lu_n;
                pending_from_raw(dw_n);
            }
        }
        if (next_snap == s0 + BL - 1) {                      // stored by the producer from the published energy
            ++snap_idx;
            const long long ns = (long long)next_snap + A.traj_every;
            next_snap = ns < s_end ? (int)ns : -1;
        }
        // ---- publish the batch, release the stage --------------------------------------------------
        if (q == 0) {
            sts64d(rs + W::O_E + 8u * (uint32_t)jg, E);
            sts32(rs + W::O_ACC + 4u * (uint32_t)jg, accb);
        }
        mbar_arrive(bar_empty + 8u * (uint32_t)stage);
        stage = stage_n; par = par_n; wb_cur = wb_n;
    }

    if (lane == 0 && n_batches) {
        atomicAdd(C.spec_stats, (unsigned long long)n_batches);
        if (n_rollbacks) atomicAdd(C.spec_stats + 1, (unsigned long long)n_rollbacks);
    }
    if (lead) C.E[c] = E;
}

template <int L, int R, bool HG>
static cudaError_t launch_ws_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    using W = WsCfg<L>;
    const SpecLayout ly = spec_layout(W::NSLOT, R, HG, T.gh4_stride);
    const long long threads = C.n_chains * L;
    const long long n_warps = (threads + 31) / 32;
    int wpb = (int)((n_warps + sms - 1) / sms);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);         // consumer warps per CTA (x 3 warps)
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    size_t smem = 0;
    for (; wpb >= 1; --wpb) {
        smem = (size_t)T.sig4_bytes + (size_t)(wpb * W::CPW) * ly.chain_bytes + (size_t)(wpb * W::NS) * W::STAGE +
               (size_t)wpb * W::EVQ + 128;
        if (smem + 1024 <= (size_t)max_optin) break;
    }
    if (wpb < 1) return cudaErrorNotSupported;
    const int grid = (int)((n_warps + wpb - 1) / wpb);
    // 4 consumer + 4 generator + 4 accounting warps: 384 threads, up to 170 registers
    cudaError_t e = cudaFuncSetAttribute(mcmc_ws_kernel<L, R, HG, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mcmc_ws_kernel<L, R, HG, 384><<<grid, wpb == 2 ? 256 : 96 * wpb, smem, st>>>(T, C, A);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_ws(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    if (!T.canonical || !T.div_safe || T.R < 1 || T.R > 8 || T.S < 1) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
    // the consumer publishes 24 sigma_index | 8 gamma_index << 16 in one word
    for (int q = 0; q < T.R; ++q)
        if (SPEC_ENT_BYTES * T.r[q].n_sigma >= 65536 || 8 * T.r[q].n_gamma >= 65536) return cudaErrorNotSupported;
#define BCP_WS(l, r) \
    case r: return hg ? launch_ws_one<l, r, true>(T, C, A, st, sms) : launch_ws_one<l, r, false>(T, C, A, st, sms);
    switch (T.R) {          // L = smallest of 2, 4, 8 that holds one lane per restraint
        BCP_WS(2, 1) BCP_WS(2, 2) BCP_WS(4, 3) BCP_WS(4, 4)
        BCP_WS(8, 5) BCP_WS(8, 6) BCP_WS(8, 7) BCP_WS(8, 8)
        default: return cudaErrorNotSupported;
    }
#undef BCP_WS
}

}  // namespace bcp
