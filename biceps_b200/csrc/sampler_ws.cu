// K5, warp-specialised speculative Metropolis kernel (the headline kernel for few chains per GPU).
//
// With 4096 chains per GPU a step costs the time ONE warp needs to issue its instruction stream in order
// (sampler_spec.cu).  Only a part of that stream is the chain's critical path -- resolve the pending step,
// evaluate the next proposal under both outcomes, commit.  Everything else is independent of the chain's
// energy and is moved to a second warp of the same CTA (same SM sub-partition: it issues in the first warp's
// dependency stalls):
//
//   consumer warp (one per 32 / L chains, L lanes per chain, one lane per restraint)
//       resolve -> evaluate -> commit, with the speculation, the exact reciprocal-table division, the log-domain
//       accept test (exact rule inside the guard band) and the rollback of a batch to a careful step-by-step
//       path of sampler_spec.cu.  It reads PRE-DECODED per-lane step words from a shared-memory ring -- the
//       shared-memory addresses of the record / gamma value to read, the sigma-entry and gamma offsets already
//       scaled to bytes, ln u -- and publishes, per batch of BL steps, the BL accept bits and the energy.
//       Indices are kept scaled (24 sigma_index, 8 gamma_index) so that an address is one add.
//   producer warp (one per consumer warp)
//       generates the Philox blocks of the steps NS batches ahead (one block per lane and batch), decodes
//       them into the per-lane words, requests the proposed states' records / gamma windows with cp.async into
//       the ring's slots, and REPLAYS the batches the consumer has finished from {proposal word, accept bit}:
//       index updates, run-length histograms (compacted per batch, one RED per event), acceptance counters,
//       state trace and the snapshots that fall on the last step of a batch (the kernel front-pads the first
//       batch so that they do whenever traj_every is a multiple of BL) -- none of this is on the consumer's
//       instruction stream.
//
// Hand-off: per consumer warp a ring of NS stages (one stage = one batch), two mbarriers per stage.
// full[stage]: 32 producer lanes arrive after their stores + 32 cp.async-completion arrivals (cp.async.mbarrier.
// arrive.noinc), so a completed phase means the ring words AND the prefetched slots have landed.  empty[stage]:
// the 32 consumer lanes arrive when the batch is done.  The consumer waits for the stage after the current one at
// the top of a batch (its last step evaluates the first step of the next batch).
// Same RNG stream, same fp64 operation order, same outputs as every other Metropolis kernel (bit for bit).
#include "spec_common.cuh"

namespace bcp {

#ifndef WS_NS
#define WS_NS 4                    // ring stages (batches in flight between the two warps)
#endif
#ifndef WS_ORDER
#define WS_ORDER 0                 // source order of a fast step: 0 resolve, evaluate; 1 evaluate, resolve; 2 evaluate's loads, resolve,
                                   // evaluate's arithmetic; 3 shuffles, evaluate's loads, the two arithmetic chains side by side (the
                                   // division's operands are tied to a shuffle result, so that ptxas cannot issue it ahead of the shuffles)
#endif
#ifndef WS_ACC64
#define WS_ACC64 0                 // fast accept test in fp64 against thresholds prepared by the producer: En < E - (ln u + margin) accepts,
                                   // En > E - (ln u - margin) rejects (one DSETP after the energy sum instead of DSUB, F2F, FADD, FSETP)
#endif
#ifndef WS_ROLLED
#define WS_ROLLED 0                // 1: the steps of a fast batch as a rolled loop (every step its own basic block: ptxas schedules the
                                   // two dependency chains of ONE step against each other instead of sinking work across steps)
#endif
#ifndef WS_LOAD_FENCE
#define WS_LOAD_FENCE 0            // 1: bar.warp.sync after a step's shared-memory loads (ptxas must not sink them below it)
#endif
#ifndef WS_TIMING_NO_REPLAY
#define WS_TIMING_NO_REPLAY 0      // 1: timing experiments only (scripts/tune_ws.sh) -- the producer skips the bookkeeping, results are WRONG
#endif

template <int L>
struct WsCfg {
    static constexpr int BL = SPEC_BL(L);        // steps per batch = per ring stage
    static constexpr int NB = BL / L;            // Philox blocks a producer lane generates per batch
    static constexpr int NS = WS_NS;
    static constexpr int CPW = 32 / L;           // chains per consumer warp
    static constexpr int NSLOT = NS * BL;        // proposal slots per chain
    // byte offsets inside a ring stage (per consumer warp)
    static constexpr int O_LW = 0;                           // [BL][32] per-LANE step words {w, recaddr, nbase, ln u}
    static constexpr int O_TH = BL * 512;                    // [BL][CPW] {ln u + margin, ln u - margin} as doubles (fast accept test)
    static constexpr int O_U = O_TH + BL * CPW * 16;         // [BL][CPW] Metropolis uniform (careful path)
    static constexpr int O_E = O_U + BL * CPW * 8;           // [CPW] energy after the batch        (consumer -> producer)
    static constexpr int O_DW = O_E + CPW * 8;               // [BL][CPW] raw proposal word (replay, careful path)
    static constexpr int O_IST = O_DW + BL * CPW * 4;        // [BL][CPW] proposed state
    static constexpr int O_WLO = O_IST + BL * CPW * 4;       // [CPW] 8 (window offset - 1) of the stage's gamma windows
    static constexpr int O_ACC = O_WLO + CPW * 4;            // [CPW] accept bits of the batch       (consumer -> producer)
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
    __shared__ __align__(8) uint64_t bars[8][2][NS];         // [consumer warp][0 full, 1 empty][stage]

    const int NW = (int)(blockDim.x >> 6);                   // consumer warps; warps NW .. 2 NW - 1 are their producers
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const bool is_producer = warp >= NW;
    const int cw = is_producer ? warp - NW : warp;
    if ((int)threadIdx.x < NW) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            mbar_init(smem_u32(&bars[threadIdx.x][0][k]), 64);
            mbar_init(smem_u32(&bars[threadIdx.x][1][k]), 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_sigma_table(reinterpret_cast<double*>(smem_raw), T.sig4, T.sig4_bytes, &tma_bar);     // has the block barrier

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
    const uint32_t bar_full = smem_u32(&bars[cw][0][0]), bar_empty = smem_u32(&bars[cw][1][0]);
    auto slot_of = [&](int stage, int j) { return slots + (uint32_t)(stage * BL + j) * (uint32_t)ly.slot_bytes; };
    auto ring_of = [&](int stage) { return ring + (uint32_t)stage * (uint32_t)W::STAGE; };

    if (is_producer) {
        // =====================================================================================
        // producer: Philox + decode + prefetch NS batches ahead, replay of the finished batches
        // =====================================================================================
        const int grp = C.group[c];
        const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
        unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
        const uint32_t sbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)T.r[qr].hist_sigma;
        const uint32_t gbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)(HG ? T.r[R - 1].hist_gamma : 0);
        const uint32_t S = (uint32_t)T.S;
        const uint32_t thresh = T.nuis_thresh;
        const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
        const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
        const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);
        const bool tracing = A.state_trace != nullptr;
        const bool traced = tracing && lead && c < A.trace_n;
        const uint32_t evq = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(NW * CPW) * (uint32_t)ly.chain_bytes +
                             (uint32_t)(NW * NS) * (uint32_t)W::STAGE + (uint32_t)cw * (uint32_t)W::EVQ;
        const uint32_t lane_lt = (1u << lane) - 1u;

        int state = C.state[c];
        int sg = C.isg[(size_t)qr * n + c];
        int gm = HG ? C.igm[(size_t)(R - 1) * n + c] : 0;
        int pg = gm;
        int since_sg = burn, since_gm = burn, since_state = burn;
        unsigned int acc_own = 0, acc_state = 0, acc_total = 0;

        uint32_t rk0[10], rk1[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            rk0[i] = seed_lo + (uint32_t)i * BCP_PHILOX_W0;
            rk1[i] = seed_hi + (uint32_t)i * BCP_PHILOX_W1;
            asm volatile("" : "+r"(rk0[i]), "+r"(rk1[i]));
        }

        // one batch: lane (chain jg, position q + m L) generates that step; then every lane writes ITS words of
        // every step of the batch
        auto generate = [&](int b, int stage) {
            const uint32_t rs = ring_of(stage);
            uint32_t g_dw[NB];
            float g_lu[NB];
            const int wofs = (gm & 1) + 4 - pg;      // smem window index of the unwrapped gamma position p: wofs + p
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
                        const double* w = gh4 + (size_t)hi * GS + (gm & ~1);
#pragma unroll
                        for (int k = 0; k < 5; ++k) cp_async16(slot + ly.slot_win + 16 * k, w + 2 * k);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < BL; ++t) {
                const uint32_t dwt = __shfl_sync(SPEC_FULL, g_dw[t / L], t % L, L);
                const float lut = __shfl_sync(SPEC_FULL, g_lu[t / L], t % L, L);
                const bool nuis = (int)dwt < 0;
                const int d1 = (int)((dwt >> fsh) & 3u) - 1, g1 = (int)((dwt >> 16) & 3u) - 1;
                const uint32_t slot = slot_of(stage, t);
                const uint32_t recaddr = nuis ? cur_rec : slot + 16 * qr;
                uint32_t nbase = recaddr;
                if (glane) nbase = nuis ? cur_row + 8u * (uint32_t)(4 + g1) : slot + (uint32_t)ly.slot_win + 8u * (uint32_t)wofs;
                const uint32_t w = (dwt & 0x80000000u) | ((uint32_t)(glane ? 8 * (g1 + 1) : 8) << 8) | (uint32_t)(EB * (d1 + 1));
                sts128u(rs + W::O_LW + 512u * (uint32_t)t + 16u * (uint32_t)lane, w, recaddr, nbase, __float_as_uint(lut));
            }
            if (q == 0) sts32(rs + W::O_WLO + 4u * (uint32_t)jg, (uint32_t)(8 * wofs - 8));
            mbar_arrive_cp_async(bar_full + 8u * (uint32_t)stage);
            mbar_arrive(bar_full + 8u * (uint32_t)stage);
        };

        // accepted state move of a replayed step (rare on large ensembles)
        auto replay_state = [&](int s, int i2) {
            ++acc_state;
            if (i2 != state) {
                if (s > burn && lead) atomicAdd(g_state_counts + state, (unsigned long long)(s - since_state));
                if (s > burn) since_state = s;
                state = i2;
            }
        };
        auto store_snapshot = [&](uint32_t rs, bool acc) {
            const size_t o = (size_t)snap_idx * n + c;
            if (lead) {
                A.traj_E[o] = lds64(rs + W::O_E + 8u * (uint32_t)jg);
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
            }
            if (owner) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = sg;
            if (HG && glane && live) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = gm;
        };
        auto next_snapshot = [&]() {
            ++snap_idx;
            const long long ns = (long long)next_snap + A.traj_every;
            next_snap = ns < s_end ? (int)ns : -1;
        };

        // the bookkeeping of a finished batch from {proposal word, accept bit}: branch-free per step, the
        // histogram events of the batch compacted into a per-warp queue and flushed with one RED each
        auto replay = [&](int b, int stage) {
            const uint32_t rs = ring_of(stage);
            const uint32_t accb = lds32(rs + W::O_ACC + 4u * (uint32_t)jg);
            const int sb = b * BL - pad;
            uint32_t evn = 0;
#pragma unroll
            for (int t = 0; t < BL; ++t) {
                const int s = sb + t;
                const uint32_t dw = lds32(rs + W::O_DW + 4u * (uint32_t)(t * CPW + jg));
                const bool acc = (accb >> t) & 1u;
                const bool rec_on = s > burn;
                const int d0 = (int)((dw >> fsh) & 3u) - 1, g0 = (int)((dw >> 16) & 3u) - 1;   // 0 for a state move
                acc_total += acc ? 1u : 0u;
                acc_own += (acc && (dw & own_mask)) ? 1u : 0u;
                int v = sg + d0;
                v += v < 0 ? nsig : 0;
                v -= v >= nsig ? nsig : 0;
                const int sg_new = acc ? v : sg;
                const bool chs = rec_on && sg_new != sg;
                uint32_t bin = sbin0 + (uint32_t)sg;
                int since = since_sg;
                bool ev = chs && owner;
                since_sg = chs ? s : since_sg;
                sg = sg_new;
                if (HG) {
                    int w = gm + g0;
                    w += w < 0 ? G : 0;
                    w -= w >= G ? G : 0;
                    const int gm_new = acc ? w : gm;
                    const bool chg = rec_on && gm_new != gm;
                    const bool evg = chg && ghist;
                    bin = evg ? gbin0 + (uint32_t)gm : bin;
                    since = evg ? since_gm : since;
                    ev = ev || evg;
                    since_gm = chg ? s : since_gm;
                    gm = gm_new;
                    pg += acc ? g0 : 0;
                }
                const uint32_t cnt = ev ? (uint32_t)(s - since) : 0u;
                const uint32_t m = __ballot_sync(SPEC_FULL, cnt != 0u);
                if (cnt != 0u) sts64u(evq + 8u * (evn + (uint32_t)__popc(m & lane_lt)), bin, cnt);
                evn += (uint32_t)__popc(m);
                if (acc && (int)dw >= 0) replay_state(s, (int)lds32(rs + W::O_IST + 4u * (uint32_t)(t * CPW + jg)));
                if (tracing) {
                    if (traced && s >= burn && s < s_end) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
                }
            }
            __syncwarp();
            for (uint32_t k = (uint32_t)lane; k < evn; k += 32) {
                const uint2 e = lds64u(evq + 8u * k);
                atomicAdd(C.param_counts + e.x, (unsigned long long)e.y);
            }
            // snapshots of this batch: the one on the last step is stored here (state == the batch's final state),
            // the others were stored by the consumer's careful path
            while (next_snap >= 0 && next_snap < sb + BL) {
                if (next_snap == sb + BL - 1) store_snapshot(rs, (accb >> (BL - 1)) & 1u);
                next_snapshot();
            }
        };

        // iteration i: replay batch i - NS (once the consumer has released its stage), then generate batch i into the
        // same stage.  Batch nbatch is generated too: the consumer's last step evaluates its first proposal (padding).
        int stage = 0, round = 0;                             // i == round * NS + stage
        for (int i = 0; i < nbatch + NS; ++i) {
            if (i >= NS) {
                mbar_wait(bar_empty + 8u * (uint32_t)stage, (uint32_t)(round - 1) & 1u);      // phase round - 1 == batch i - NS
#if !WS_TIMING_NO_REPLAY
                replay(i - NS, stage);
#endif
                __syncwarp();
            }
            if (i <= nbatch) generate(i, stage);
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
            if (s_end > since_sg) atomicAdd(C.param_counts + sbin0 + sg, (unsigned long long)(s_end - since_sg));
            C.isg[(size_t)q * n + c] = sg;
            C.sep[(size_t)q * n + c] += acc_own;
        }
        if (ghist) {
            if (s_end > since_gm) atomicAdd(C.param_counts + gbin0 + gm, (unsigned long long)(s_end - since_gm));
            C.igm[(size_t)(R - 1) * n + c] = gm;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
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
    auto bcast_d = [&](double v, int src) { return __shfl_sync(SPEC_FULL, v, src, L); };
    const uint32_t zero = (uint32_t)((unsigned long long)A.n_snap >> 63);      // 0 at run time, opaque to the compiler (WS_ORDER 3)

    int state = C.state[c];
    double E = C.E[c];
    int sg24 = EB * C.isg[(size_t)qr * n + c];               // 24 x this lane's sigma index
    int gm8 = glane ? 8 * C.igm[(size_t)(R - 1) * n + c] : 0;  // 8 x gamma index, in the gamma lane only (0 elsewhere)
    int pg8 = gm8;                                           // 8 x unwrapped gamma position (window bookkeeping)
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
    struct Cand { double xA, xB; int a1, g1x; bool miss; };
    struct EvalIn { double nlA, dvA, yA, nlB, dvB, yB, numA, numB, ref, Cn; int a1, g1x; bool miss; };
    auto evaluate_load = [&](const uint4& lw, int wlo8, int a0_, int g0x_) {
        EvalIn k;
        const bool nuis_n = (int)lw.x < 0;
        const uint32_t ab = lw.x & 0xffu;
        k.a1 = (int)ab - EB;
        k.g1x = (int)((lw.x >> 8) & 0xffu) - 8;
        const uint32_t eA = entm + (uint32_t)sg24 + ab;      // unwrapped: entries carry a halo of 2
        const uint32_t eB = eA + (uint32_t)a0_;
        k.nlA = lds64(eA); k.dvA = lds64(eA + 8); k.yA = lds64(eA + 16);
        k.nlB = lds64(eB); k.dvB = lds64(eB + 8); k.yB = lds64(eB + 16);
        const uint32_t recaddr = lw.y;
        const double2 rv = lds128(recaddr);
        k.numA = rv.x; k.numB = rv.x; k.ref = rv.y;
        k.miss = false;
        if (HG) {
            // the other lanes re-read their own sse word: nbase == recaddr, gm8 == pg8 == g0x == wlo8 == 0 there
            k.miss = !nuis_n && (uint32_t)(pg8 + wlo8) > 56u;    // window index +- 1 must stay inside the 10-wide window
            uint32_t na = lw.z + (uint32_t)(nuis_n ? gm8 : pg8);
            na = k.miss ? recaddr : na;
            k.numA = lds64(na);
            k.numB = lds64(na + (uint32_t)g0x_);
        }
        k.Cn = Cst;
        if (!nuis_n && q == 0) k.Cn = __dadd_rn(lds64(recaddr + 16 * R), logZ);
        return k;
    };
    auto evaluate_math = [&](const EvalIn& in) {
        Cand k;
        k.a1 = in.a1; k.g1x = in.g1x; k.miss = in.miss;
        double t = __dadd_rn(in.nlA, div_markstein(in.numA, in.dvA, in.yA));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, in.ref);
        k.xA = __dadd_rn(t, in.Cn);
        t = __dadd_rn(in.nlB, div_markstein(in.numB, in.dvB, in.yB));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, in.ref);
        k.xB = __dadd_rn(t, in.Cn);
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
            pg8 += acc ? g0x : 0;
        }
    };
    auto pending_from_raw = [&](uint32_t dn) {
        a0 = EB * ((int)((dn >> fsh) & 3u) - 1);
        g0x = glane ? 8 * ((int)((dn >> 16) & 3u) - 1) : 0;
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
    auto wlo_of = [&](uint32_t rs) { return glane ? (int)lds32(rs + W::O_WLO + 4u * (uint32_t)jg) : 0; };
    int wlo_cur = wlo_of(ring_of(0));

    int stage = 0;
    uint32_t par = 0;                                         // parity of the current stage's full phase
    for (int b = 0; b < nbatch; ++b) {
        const int s0 = b * BL - pad;
        const int stage_n = stage + 1 == NS ? 0 : stage + 1;
        const uint32_t par_n = stage_n == 0 ? par ^ 1u : par;
        const uint32_t rs = ring_of(stage), rn = ring_of(stage_n);
        mbar_wait(bar_full + 8u * (uint32_t)stage_n, par_n);  // the last step of this batch evaluates the next batch's first step
        const int wlo_n = wlo_of(rn);
        uint32_t accb = 0;

        // ---- fast batch: BL steps as ONE basic block on registers only (shared memory is read, never written);
        //      anything rare only raises a flag and the batch is then redone carefully ------------------------
        bool careful = (unsigned)(next_snap - s0) < (unsigned)(BL - 1);   // a snapshot before the batch's last step
        if (!careful) {
            // the step words of the proposals evaluated in this batch: positions 1 .. BL-1 and the next batch's first
#if WS_ROLLED
            uint32_t lw_addr = rs + 512u + lw_lane;          // words of position 1; the last step reads the next stage's position 0
#else
            uint4 lw[BL];
#pragma unroll
            for (int j = 0; j < BL; ++j) lw[j] = lds128u((j + 1 == BL ? rn : rs + 512u * (uint32_t)(j + 1)) + lw_lane);
#endif
            const double sv_E = E, sv_x = x_prop;
            const int sv_sg = sg24, sv_gm = gm8, sv_pg = pg8, sv_a0 = a0, sv_g0 = g0x;
            const bool sv_nuis = nuis_s;
            const float sv_lu = lu_s;
            bool rare = false;
#if WS_ROLLED
#pragma unroll 1
            for (int j = 0; j < BL; ++j) {
                const bool last = j + 1 == BL;
                const uint4 lwj = lds128u(last ? rn + lw_lane : lw_addr);
                lw_addr += 512u;
                const int wlo_j = last ? wlo_n : wlo_cur;
#else
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                const uint4 lwj = lw[j];
                const int wlo_j = j + 1 == BL ? wlo_n : wlo_cur;
#endif
                // (1) resolve the pending step: energy in the reference's summation order, accept in the log domain.
                // (2) contributions to the next step under both outcomes of the pending one: they depend on the commit of
                //     the step BEFORE the pending one only, so the two dependency chains (shuffles -> 3 adds -> accept;
                //     loads -> division -> 4 adds) can run side by side.  ptxas serialises them when left alone (it issues
                //     the division ahead of the shuffles: 460 cycles per step); WS_ORDER 3 issues the shuffles first and ties
                //     the division's numerators to the last shuffle's result (x ^ (v & 0), one LOP3 each).
#if WS_ORDER == 1
                const Cand k = evaluate_math(evaluate_load(lwj, wlo_j, a0, g0x));
#elif WS_ORDER == 2
                const EvalIn ein = evaluate_load(lwj, wlo_j, a0, g0x);
#endif
                double xv[R];
#pragma unroll
                for (int r = 0; r < R; ++r) xv[r] = bcast_d(x_prop, r);
#if WS_ACC64
                // thresholds of the pending step (position j of this stage) against the committed energy
                const double2 th = lds128(rs + W::O_TH + 16u * (uint32_t)jg + (uint32_t)(16 * CPW) * (uint32_t)j);
                const double Ahi = __dsub_rn(E, th.x), Alo = __dsub_rn(E, th.y);
#endif
#if WS_ORDER == 3
                EvalIn ein = evaluate_load(lwj, wlo_j, a0, g0x);
#if WS_LOAD_FENCE
                asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
#endif
                {
                    const uint32_t tie = (uint32_t)__double2loint(xv[R - 1]) & zero;
                    ein.numA = __hiloint2double(__double2hiint(ein.numA), (int)((uint32_t)__double2loint(ein.numA) ^ tie));
                    ein.numB = __hiloint2double(__double2hiint(ein.numB), (int)((uint32_t)__double2loint(ein.numB) ^ tie));
                }
#endif
                double En = xv[0];
#pragma unroll
                for (int r = 1; r < R; ++r) En = __dadd_rn(En, xv[r]);
#if WS_ACC64
                const bool acc = En < Ahi;
                const bool unsure = !acc && !(En > Alo);                 // guard band, also NaN
#else
                const double d = __dsub_rn(E, En);
                const float tt = (float)d - lu_s;
                const bool acc = tt > SPEC_MARGIN;
                const bool unsure = !acc && !(tt < -SPEC_MARGIN);        // guard band, also NaN
#endif
#if WS_ORDER == 0
                const Cand k = evaluate_math(evaluate_load(lwj, wlo_j, a0, g0x));
#elif WS_ORDER >= 2
                const Cand k = evaluate_math(ein);
#endif
                rare = rare || unsure || (acc && !nuis_s) || k.miss;
                // (3) commit
                commit(acc, En);
                accb |= acc ? (1u << j) : 0u;
                x_prop = acc ? k.xB : k.xA;
                nuis_s = (int)lwj.x < 0; lu_s = __uint_as_float(lwj.w); a0 = k.a1; g0x = k.g1x;
            }
            if (__any_sync(SPEC_FULL, rare)) {               // roll the batch back
                E = sv_E; x_prop = sv_x; sg24 = sv_sg; gm8 = sv_gm; pg8 = sv_pg; a0 = sv_a0; g0x = sv_g0;
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
        stage = stage_n; par = par_n; wlo_cur = wlo_n;
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
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
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
    // up to 4 consumer + 4 producer warps: no register cap (255); the 8 + 8 shape (two waves) is capped at 128
    if (wpb <= 4) {
        cudaError_t e = cudaFuncSetAttribute(mcmc_ws_kernel<L, R, HG, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mcmc_ws_kernel<L, R, HG, 256><<<grid, 64 * wpb, smem, st>>>(T, C, A);
    } else {
        cudaError_t e = cudaFuncSetAttribute(mcmc_ws_kernel<L, R, HG, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mcmc_ws_kernel<L, R, HG, 512><<<grid, 64 * wpb, smem, st>>>(T, C, A);
    }
    return cudaGetLastError();
}

cudaError_t launch_mcmc_ws(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    if (!T.canonical || !T.div_safe || T.R < 1 || T.R > 8 || T.S < 1) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
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
