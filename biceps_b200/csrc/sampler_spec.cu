// K5, speculative cooperative Metropolis kernel: L lanes per chain, one lane per restraint, the
// candidate energy of step s+1 prepared while step s is being resolved.
//
// With few chains per GPU (the C3 configuration runs 4096 = one warp per SM sub-partition at L = 4)
// a step costs the time ONE warp needs to issue its instruction stream in order.  The design therefore
// minimises (i) the dependent latency between two consecutive accept decisions, (ii) the number of
// instructions per step and (iii) the number of basic blocks per step, so that ptxas can interleave the
// independent work (measured latencies: scripts/microbench.cu -> profiles/r01_microbench.json).
//
//   * Proposals depend only on the counter-based RNG, so while step s is resolved every lane
//     evaluates ITS restraint term of step s+1's proposal under both outcomes of step s
//     (H0: s rejected, H1: s accepted; they differ only in the lane that step s moves).  The accept
//     decision then selects one of the two.  Dependent path per step: select -> R shuffles ->
//     R-1 fp64 adds (reference order; lane 0 has pre-added energy + logZ) -> subtract -> compare.
//   * sse / (2 sigma^2) uses a correctly rounded reciprocal table and Markstein's correction
//     (5 dependent FMA-class operations instead of the ~124-cycle IEEE division sequence); the result
//     is bit-identical to IEEE division (see div_markstein).
//   * The Metropolis test u < exp(d) is decided as d - ln(u) > +-2^-13 with ln(u) computed in fp32 by
//     the lane that generated the step's Philox block, ahead of time; inside the guard band the
//     specified exact rule (exp_det.cuh) is evaluated, so the decision is always the specified one.
//   * All state-dependent words of a chain live in shared memory: the current state's record and its
//     whole gamma row (periodic halo, so index +-2 needs no wrap), and, for the state moves of the
//     next batches, the proposed state's record + a 10-wide gamma window, fetched with cp.async
//     two batches ahead by the lane that generated that step's Philox block (three batches of slots are
//     alive: the current one, the next one, the one being generated).  Chain blocks are laid out 16 B apart
//     modulo the 128-B bank period, so the 8 chains of a warp never collide.
//   * Lane q of a group generates the Philox blocks of positions q, q + L, ... of a batch of BL steps
//     (BL = 4 for L <= 4, 8 for L = 8; 8 was measured slower at L = 4); the 10 rounds of a block are spread
//     over L steps of the batch two batches earlier (software pipelined by hand).
//   * The sigma vectors {Ndof ln sigma, 2 sigma^2, 1/(2 sigma^2)} of all restraints are staged once per
//     CTA by a 1-D TMA bulk copy (UBLKCP).
//   * One step is ONE basic block.  ptxas turns every predicated RED into a divergent branch, so the
//     run-length histogram flushes are not issued in the step: a lane appends {bin, count} to a per-warp
//     queue in shared memory (ballot + popc, predicated STS) and the queue is drained by one RED per
//     batch.  Everything rare (guard band of the accept test, an accepted state move = record copy +
//     gamma row reload, a gamma index that drifted out of a prefetched window, a trajectory snapshot /
//     state-trace store) sits behind a single warp-uniform branch that re-does the step's commit exactly.
// Same RNG stream, same fp64 operation order and the same outputs as the other kernels (parity with the
// oracle is bit for bit); canonical restraint layout only (tables.cuh).
#include "spec_common.cuh"

namespace bcp {

#ifndef SPEC_REGCACHE
#define SPEC_REGCACHE 0            // 1: the words a lane's term depends on (sigma entry of its index, its record, the gamma value) are kept
                                   // in registers and shared memory is read -- with predicated loads -- only by the lanes whose inputs
                                   // change in the evaluated proposal: 4 warps per SM otherwise keep the SM's shared-memory pipe ~70 %
                                   // busy (bank conflicts of 32 random addresses per load included), which is what stretches every
                                   // dependent load of a step
#endif


template <int L, int R, bool HG, bool TRACE>
__global__ void __launch_bounds__(256, SPEC_MINBLOCKS)
mcmc_spec_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                 const __grid_constant__ LaunchArgs A) {
    constexpr int D = SpecCfg<L>::D;
    constexpr int BL = SpecCfg<L>::BL;
    constexpr int NB = SpecCfg<L>::NB;
    constexpr int NSLOT = SpecCfg<L>::NSLOT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(reinterpret_cast<double*>(smem_raw), T.sig4, T.sig4_bytes, &mbar);

    const long long n = C.n_chains;
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(threadIdx.x & (L - 1));             // lane within the chain's group
    long long c = gthread / L;
    const bool live = c < n;                                // dead groups shadow the last chain, no stores
    if (!live) c = n - 1;
    const int qr = q < R ? q : 0;                           // restraint this lane evaluates (idle lanes: 0)
    const bool owner = live && q < R;
    const bool lead = live && q == 0;
    const bool glane = HG && q == R - 1;                    // lane of the gamma restraint
    const uint32_t own_mask = q < R ? (1u << (20 + q)) : 0u;   // this lane's restraint moves
    const int fsh = 2 * qr;                                  // position of this lane's dsigma field
    constexpr int GL = R < L ? R : 0;                        // lane that keeps the gamma histogram (an idle lane if any)
    const bool ghist = HG && live && q == GL;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double* __restrict__ rec = T.rec;
    const double* __restrict__ gh4 = HG ? T.gsse_halo4 : nullptr;
    const int G = HG ? T.r[R - 1].n_gamma : 1;
    const int GS = HG ? T.gh4_stride : 0;
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    const uint32_t sbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)T.r[qr].hist_sigma;      // bins of C.param_counts
    const uint32_t gbin0 = (uint32_t)grp * (uint32_t)T.HB + (uint32_t)(HG ? T.r[R - 1].hist_gamma : 0);
    const uint32_t S = (uint32_t)T.S;
    const uint32_t thresh = T.nuis_thresh;
    const int s_end = (int)(A.s_end - A.s_begin);
    const long long burn_l = A.burn - A.s_begin;
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
    const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
    const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);

    const SpecLayout ly = spec_layout(NSLOT, R, HG, GS);
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t chain_base = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(threadIdx.x / L) * (uint32_t)ly.chain_bytes;
    const uint32_t cur_rec = chain_base + ly.cur_rec + 16 * qr;
    const uint32_t cur_row = chain_base + ly.cur_row;
    const uint32_t slots = chain_base + ly.slot0;
    // per-warp queue of histogram events {bin, count}: at most 2 per chain and step, 128 per batch
    const uint32_t evq = smem0 + (uint32_t)T.sig4_bytes + (uint32_t)(blockDim.x / L) * (uint32_t)ly.chain_bytes +
                         (uint32_t)(threadIdx.x >> 5) * 1024u;
    const uint32_t lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    const int nsig = T.r[qr].n_sigma;
    const uint32_t ent_base = smem0 + (uint32_t)SPEC_ENT_BYTES * (uint32_t)(T.ent_off[qr] + 2);   // entry of sigma index 0
    const double cnorm = T.r[qr].cnorm;

    // ---- chain state: E and state replicated in every lane of the group --------------------------
    int state = C.state[c];
    double E = C.E[c];
    int sg = C.isg[(size_t)qr * n + c];
    int gm = HG ? C.igm[(size_t)(R - 1) * n + c] : 0;        // replicated in every lane of the group
    int pg = gm;                                             // unwrapped gamma position (window bookkeeping)
    // lane 0 adds energy + logZ to its term (reference order (C + t_0) + t_1 ...); x + (-0.0) == x bitwise
    double Cst = q == 0 ? __dadd_rn(energy[state], logZ) : -0.0;
    int since_sg = burn, since_gm = burn, since_state = burn;
    unsigned int acc_own = 0, acc_state = 0, acc_total = 0;
    unsigned int n_batches = 0, n_rollbacks = 0;             // per warp, for the host's kernel choice

    if (q < R) sts128(cur_rec, *reinterpret_cast<const double2*>(rec + ((size_t)state * R + q) * 2));
    if (HG) {
        const double2* src = reinterpret_cast<const double2*>(gh4 + (size_t)state * GS);
        for (int k = q; k < GS / 2; k += L) sts128(cur_row + 16 * k, src[k]);
    }

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }
    // the state only changes in the careful path, so a fast batch traces L copies of it at its end (TRACE
    // kernels only: the extra code costs the untraced kernel ~6 % through ptxas' schedule)
    const bool traced = TRACE && lead && c < A.trace_n;
    int next_ev = next_snap;                                 // next step that must take the careful path

    // ---- Philox block of one step, generated by the lane whose index is the step's batch position ----
    // the 20 round keys are the same for every block of the launch: kept in registers (made opaque, so that
    // the compiler does not re-derive them from the seed in every round)
    uint32_t rk0[10], rk1[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        rk0[i] = seed_lo + (uint32_t)i * BCP_PHILOX_W0;
        rk1[i] = seed_hi + (uint32_t)i * BCP_PHILOX_W1;
        asm volatile("" : "+r"(rk0[i]), "+r"(rk1[i]));
    }
    uint32_t pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0;
    auto philox_begin = [&](int s) {
        const unsigned long long ctr = ctr0 + (unsigned long long)s;
        pc0 = (uint32_t)ctr; pc1 = (uint32_t)(ctr >> 32); pc2 = ch_lo; pc3 = ch_hi;
    };
    auto philox_round = [&](int i) {
        const uint64_t p0 = (uint64_t)BCP_PHILOX_M0 * pc0;
        const uint64_t p1 = (uint64_t)BCP_PHILOX_M1 * pc2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ pc1 ^ rk0[i];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ pc3 ^ rk1[i];
        pc1 = (uint32_t)p1; pc3 = (uint32_t)p0; pc0 = n0; pc2 = n2;
    };
    auto philox_finish = [&](int s) {
        const bool nuis = pc0 < thresh;
        const uint64_t pr = (uint64_t)pc1 * (nuis ? (uint32_t)R : S);
        const uint32_t hi = (uint32_t)(pr >> 32);
        const uint64_t z1 = (uint64_t)(uint32_t)pr * 3u;
        const uint64_t z2 = (uint64_t)(uint32_t)z1 * 3u;
        const uint32_t a = (uint32_t)(z1 >> 32);
        const uint32_t b = (HG && hi == (uint32_t)(R - 1)) ? (uint32_t)(z2 >> 32) : 1u;
        SpecGen g;
#if SPEC_FINISH_BRANCHFREE
        uint32_t dwn = (0x80000000u | SPEC_DW_IDLE | (0x00100000u << hi)) ^ ((a ^ 1u) << (2 * hi)) ^ ((b ^ 1u) << 16);
        asm volatile("" : "+r"(dwn));
        g.dw = nuis ? dwn : SPEC_DW_IDLE;
#else
        g.dw = nuis ? ((0x80000000u | SPEC_DW_IDLE | (0x00100000u << hi)) ^ ((a ^ 1u) << (2 * hi)) ^ ((b ^ 1u) << 16))
                    : SPEC_DW_IDLE;
#endif
        g.istate = hi;
        g.u = spec_u2(pc2, pc3);
        // steps past the end of the launch are executed as padding of the last batch: never accepted
        g.lu = s < s_end ? __logf((float)g.u) : __int_as_float(0x7f800000);
        return g;
    };
    // cp.async of the proposed state's words for a generated state move into the step's slot
    auto slot_of = [&](int sbatch, int j) { return slots + (uint32_t)(sbatch * BL + j) * (uint32_t)ly.slot_bytes; };
    auto prefetch = [&](const SpecGen& g, uint32_t slot, int gm_now) {
        if (!((int)g.dw < 0)) {
            const uint32_t i2 = g.istate;
            const double* r2 = rec + (size_t)i2 * (2 * R);
#pragma unroll
            for (int k = 0; k < R; ++k) cp_async16(slot + 16 * k, r2 + 2 * k);
            cp_async8(slot + ly.slot_c, energy + i2);
            if (HG) {
                const double* w = gh4 + (size_t)i2 * GS + (gm_now & ~1);
#pragma unroll
                for (int k = 0; k < 5; ++k) cp_async16(slot + ly.slot_win + 16 * k, w + 2 * k);
            }
        }
    };
    auto bcast_u = [&](uint32_t v, int src) { return __shfl_sync(SPEC_FULL, v, src, L); };
    auto bcast_f = [&](float v, int src) { return __shfl_sync(SPEC_FULL, v, src, L); };
    auto bcast_d = [&](double v, int src) { return __shfl_sync(SPEC_FULL, v, src, L); };
    // window offset of a batch prefetched now: smem window index of unwrapped position p is wofs + p
    auto window_ofs = [&]() { return (gm & 1) + 4 - pg; };

    // ---- this lane's contribution to the energy of step `sn` (word dn): its restraint term, plus
    //      energy + logZ in lane 0.  A: on the committed state, B: on the committed state shifted by the
    //      pending step (d0, g0).  Window misses are only flagged here (out-of-line path re-evaluates).
    struct Cand { double xA, xB; int d1, g1; bool own, miss; };
    auto evaluate = [&](uint32_t dn, uint32_t slot, int wofs, int d0_, int g0_) {
        Cand k;
        const bool nuis_n = (int)dn < 0;
        k.own = (dn & own_mask) != 0u;
        k.d1 = (int)((dn >> fsh) & 3u) - 1;
        k.g1 = (int)((dn >> 16) & 3u) - 1;
        const uint32_t eA = ent_base + (uint32_t)(SPEC_ENT_BYTES * (sg + k.d1));   // unwrapped: entries carry a halo of 2
        const uint32_t eB = eA + (uint32_t)(SPEC_ENT_BYTES * d0_);
        const double nlA = lds64(eA), dvA = lds64(eA + 8), yA = lds64(eA + 16);
        const double nlB = lds64(eB), dvB = lds64(eB + 8), yB = lds64(eB + 16);
        const uint32_t recaddr = nuis_n ? cur_rec : slot + 16 * qr;
        const double2 rv = lds128(recaddr);
        double numA = rv.x, numB = rv.x;
        k.miss = false;
        if (HG) {
            // branch-free: the other lanes re-read their own sse word (the address of rv.x, stride 0)
            const int wi = wofs + pg;                        // window index of the current position (state move)
            const bool out = (unsigned)(wi - 1) > 7u;        // wi +- 1 must stay inside the 10-wide window
            const int ia = nuis_n ? gm + 4 + k.g1 : (out ? 4 : wi);
            const uint32_t na = glane ? (nuis_n ? cur_row : slot + ly.slot_win) + 8u * (uint32_t)ia : recaddr;
#if SPEC_PRED_GAMMA
            numA = lds64_if(glane, na, numA);
            numB = lds64_if(glane, na + (uint32_t)(8 * g0_), numB);
#else
            numA = lds64(na);
            numB = lds64(na + (uint32_t)(glane ? 8 * g0_ : 0));
#endif
            k.miss = glane && !nuis_n && out;
        }
        double Cn = Cst;
        if (!nuis_n && q == 0) Cn = __dadd_rn(lds64(slot + ly.slot_c), logZ);
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
#if SPEC_REGCACHE
    // register copies (valid inside fast batches; reloaded after every careful batch):
    //   c_*  entry / record / gamma value of the committed indices,   p_*  of the committed indices shifted by the
    //   pending step (d0, g0) -- the pending step's own evaluation loaded them one step earlier
    double c_nl = 0.0, c_dv = 1.0, c_y = 1.0, c_sse = 0.0, c_ref = 0.0, c_num = 0.0;
    double p_nl = 0.0, p_dv = 1.0, p_y = 1.0, p_num = 0.0;
    double a_nl, a_dv, a_y, a_num, b_nl, b_dv, b_y, b_num;      // entries of the two candidates of the step being evaluated
    auto reload_cache = [&](int d0_, int g0_) {
        const uint32_t e0 = ent_base + (uint32_t)(SPEC_ENT_BYTES * sg), e1 = e0 + (uint32_t)(SPEC_ENT_BYTES * d0_);
        c_nl = lds64(e0); c_dv = lds64(e0 + 8); c_y = lds64(e0 + 16);
        p_nl = lds64(e1); p_dv = lds64(e1 + 8); p_y = lds64(e1 + 16);
        const double2 rv = lds128(cur_rec);
        c_sse = rv.x; c_ref = rv.y;
        if (HG) {
            c_num = glane ? lds64(cur_row + 8u * (uint32_t)(gm + 4)) : 0.0;
            p_num = glane ? lds64(cur_row + 8u * (uint32_t)(gm + 4 + g0_)) : 0.0;
        }
    };
    auto evaluate_rc = [&](uint32_t dn, uint32_t slot, int wofs, int d0_, int g0_) {
        Cand k;
        const bool nuis_n = (int)dn < 0;
        k.own = (dn & own_mask) != 0u;
        k.d1 = (int)((dn >> fsh) & 3u) - 1;
        k.g1 = (int)((dn >> 16) & 3u) - 1;
        // sigma entries: A at sg + d1, B at sg + d0 + d1; every case but "d != 0 and not the pending index" is in registers
        const uint32_t eA = ent_base + (uint32_t)(SPEC_ENT_BYTES * (sg + k.d1));
        const uint32_t eB = eA + (uint32_t)(SPEC_ENT_BYTES * d0_);
        const bool ldA = k.d1 != 0;
        const bool ldB = d0_ != 0 && k.d1 == d0_;                  // sg + 2 d0: the only B index that is neither A's, the pending nor the current one
        a_nl = lds64_if(ldA, eA, c_nl); a_dv = lds64_if(ldA, eA + 8, c_dv); a_y = lds64_if(ldA, eA + 16, c_y);
        b_nl = lds64_if(ldB, eB, a_nl); b_dv = lds64_if(ldB, eB + 8, a_dv); b_y = lds64_if(ldB, eB + 16, a_y);
        if (d0_ != 0 && k.d1 == 0) { b_nl = p_nl; b_dv = p_dv; b_y = p_y; }             // B == the pending index
        if (d0_ != 0 && k.d1 == -d0_) { b_nl = c_nl; b_dv = c_dv; b_y = c_y; }          // B == the current index
        // record: the current state's (registers) or the proposed state's (slot)
        double sse = c_sse, ref = c_ref;
        {
            double2 rv = make_double2(c_sse, c_ref);
            rv = lds128_if(!nuis_n, slot + 16 * qr, rv);
            sse = rv.x; ref = rv.y;
        }
        a_num = sse; b_num = sse;
        k.miss = false;
        if (HG) {
            const int wi = wofs + pg;                        // window index of the current position (state move)
            const bool out = (unsigned)(wi - 1) > 7u;        // wi +- 1 must stay inside the 10-wide window
            const bool gl_n = glane && nuis_n, gl_s = glane && !nuis_n;
            // nuisance proposal: values of the current row at gm + g1 (A) and gm + g0 + g1 (B)
            const uint32_t rA = cur_row + 8u * (uint32_t)(gm + 4 + k.g1);
            const bool ldgA = gl_n && k.g1 != 0;
            const bool ldgB = gl_n && g0_ != 0 && k.g1 == g0_;
            double nA = glane ? c_num : sse;
            nA = lds64_if(ldgA, rA, nA);
            double nB = nA;
            nB = lds64_if(ldgB, rA + (uint32_t)(8 * g0_), nB);
            if (gl_n && g0_ != 0 && k.g1 == 0) nB = p_num;
            if (gl_n && g0_ != 0 && k.g1 == -g0_) nB = c_num;
            // state-move proposal: the slot's window at the current gamma position (A) and shifted by g0 (B)
            const uint32_t wA = slot + ly.slot_win + 8u * (uint32_t)(out ? 4 : wi);
            nA = lds64_if(gl_s, wA, nA);
            nB = gl_s ? nA : nB;
            nB = lds64_if(gl_s && g0_ != 0, wA + (uint32_t)(8 * g0_), nB);
            a_num = nA; b_num = nB;
            k.miss = gl_s && out;
        }
        double Cn = Cst;
        if (!nuis_n && q == 0) Cn = __dadd_rn(lds64(slot + ly.slot_c), logZ);
        double t = __dadd_rn(a_nl, div_markstein(a_num, a_dv, a_y));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, ref);
        k.xA = __dadd_rn(t, Cn);
        t = __dadd_rn(b_nl, div_markstein(b_num, b_dv, b_y));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, ref);
        k.xB = __dadd_rn(t, Cn);
        return k;
    };
    // after the decision on the pending step: the committed copies follow an accepted move (state moves roll the batch
    // back), the candidate whose history came true becomes the new pending copy.  nuis_n: the evaluated step is a nuisance move
    auto advance_cache = [&](bool acc, bool nuis_n) {
        c_nl = acc ? p_nl : c_nl; c_dv = acc ? p_dv : c_dv; c_y = acc ? p_y : c_y;
        p_nl = acc ? b_nl : a_nl; p_dv = acc ? b_dv : a_dv; p_y = acc ? b_y : a_y;
        if (HG) {
            c_num = acc ? p_num : c_num;
            // the pending copy of the gamma value only matters for a nuisance pending step (an accepted state move rolls back)
            p_num = nuis_n ? (acc ? b_num : a_num) : c_num;
        }
    };
#endif
    // the same on the committed state only; the gamma value of a proposed state (i2) is read from global memory
    auto evaluate_exact = [&](uint32_t dn, uint32_t slot, uint32_t i2) {
        const bool nuis_n = (int)dn < 0;
        const int d1 = (int)((dn >> fsh) & 3u) - 1, g1 = (int)((dn >> 16) & 3u) - 1;
        const uint32_t eA = ent_base + (uint32_t)(SPEC_ENT_BYTES * (sg + d1));
        const double nlA = lds64(eA), dvA = lds64(eA + 8), yA = lds64(eA + 16);
        const double2 rv = lds128(nuis_n ? cur_rec : slot + 16 * qr);
        double num = rv.x;
        if (HG) {
            if (glane) num = nuis_n ? lds64(cur_row + 8u * (uint32_t)(gm + 4 + g1)) : __ldg(gh4 + (size_t)i2 * GS + (gm + 4));
        }
        double Cn = Cst;
        if (!nuis_n && q == 0) Cn = __dadd_rn(lds64(slot + ly.slot_c), logZ);
        double t = __dadd_rn(nlA, div_markstein(num, dvA, yA));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, rv.y);
        return __dadd_rn(t, Cn);
    };

    // ---- pending step (to be resolved next) and the commit of a decision ---------------------------
    uint32_t dw_s;
    float lu_s;
    double x_prop;
    int d0, g0;          // this lane's sigma shift and the group's gamma shift if the pending step is accepted
    bool own_s;
    // commit of step s; returns this lane's histogram event {bin, run length} (count 0 = none)
    auto commit = [&](bool acc, int s, double En, uint32_t& ev_bin, uint32_t& ev_cnt) {
        const bool rec_on = s > burn;
        E = acc ? En : E;
        acc_total += acc ? 1u : 0u;
        acc_own += (acc && own_s) ? 1u : 0u;
        int v = sg + d0;                                      // d0 == 0 unless this lane's restraint moves
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
            int w = gm + g0;                                  // replicated: every lane follows the gamma index
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
        ev_bin = bin;
        ev_cnt = ev ? (uint32_t)(s - since) : 0u;
    };

    // ---- prologue: batches 0 .. D-1 generated and prefetched, contribution of step 0 ----------------
    // lane q generates positions q, q + L, ... of a batch: gen[b][m] = step (batch b) * BL + m * L + q
    SpecGen gen[D][NB];
    int wofs[D + 1];
    int sb0 = 0;                                             // slot batch of the current batch (0 .. NBATCH-1)
    {
#pragma unroll
        for (int b = 0; b < D; ++b) {
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const int sp = b * BL + m * L + q;
                philox_begin(sp);
#pragma unroll
                for (int r = 0; r < 10; ++r) philox_round(r);
                gen[b][m] = philox_finish(sp);
                prefetch(gen[b][m], slot_of(b, m * L + q), gm);
            }
            wofs[b] = window_ofs();
        }
        wofs[D] = wofs[0];
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        dw_s = bcast_u(gen[0][0].dw, 0);
        lu_s = bcast_f(gen[0][0].lu, 0);
        own_s = (dw_s & own_mask) != 0u;
        d0 = (int)((dw_s >> fsh) & 3u) - 1;
        g0 = (int)((dw_s >> 16) & 3u) - 1;
        x_prop = evaluate_exact(dw_s, slot_of(0, 0), bcast_u(gen[0][0].istate, 0));
    }

#if SPEC_REGCACHE
    bool cache_stale = true;
#endif
    for (int s0 = 0; s0 < s_end; s0 += BL) {
        const int sb1 = sb0 + 1 == SpecCfg<L>::NBATCH ? 0 : sb0 + 1;         // slot batch of the next batch
        const int sb2 = sb0 + D >= SpecCfg<L>::NBATCH ? sb0 + D - SpecCfg<L>::NBATCH : sb0 + D;   // ... of the batch generated now
        if (D >= 3) {            // the next batch's proposals (requested two batches ago) are read by this batch's last step
            cp_async_wait<1>();
            __syncwarp();
        }
        uint32_t ev_bin[BL], ev_cnt[BL];
#pragma unroll
        for (int j = 0; j < BL; ++j) { ev_bin[j] = 0; ev_cnt[j] = 0; }
        SpecGen gnew[NB];

        // ---- fast batch: BL steps as ONE basic block on registers only (shared memory is read, never
        //      written); anything rare only raises a flag and the batch is then redone carefully -------
        bool careful = (unsigned)(next_ev - s0) < (unsigned)BL;         // a snapshot falls into it
        if (!careful) {
#if SPEC_REGCACHE
            if (cache_stale) {                                           // (warp-uniform) after a careful batch
                reload_cache(d0, g0);
                cache_stale = false;
            }
#endif
            const double sv_E = E, sv_x = x_prop;
            const int sv_sg = sg, sv_gm = gm, sv_pg = pg, sv_ssg = since_sg, sv_sgm = since_gm, sv_d0 = d0, sv_g0 = g0;
            const unsigned int sv_at = acc_total, sv_ao = acc_own;
            const uint32_t sv_dw = dw_s;
            const float sv_lu = lu_s;
            const bool sv_own = own_s;
            bool rare = false;
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                const int s = s0 + j;
                // Philox: block m of the batch generated now runs its 10 rounds during steps m L .. m L + L - 1
                if (j % L == 0) philox_begin(s0 + D * BL + j + q);
#pragma unroll
                for (int r = 0; r < 10; ++r)
                    if (r * L / 10 == j % L) philox_round(r);
                if (j % L == L - 1) gnew[j / L] = philox_finish(s0 + D * BL + (j / L) * L + q);
                // the proposals of the next batch (requested one batch ago) are first read by the last step
                if (D < 3 && j == BL - 1) {
                    cp_async_wait<0>();
                    __syncwarp();
                }
                const bool nb = j + 1 == BL;
                const int jn = nb ? 0 : j + 1;
                const uint32_t dw_n = bcast_u(nb ? gen[1][0].dw : gen[0][jn / L].dw, jn % L);
                const float lu_n = bcast_f(nb ? gen[1][0].lu : gen[0][jn / L].lu, jn % L);
#if SPEC_EVAL_FIRST && !SPEC_REGCACHE
                const Cand k = evaluate(dw_n, slot_of(nb ? sb1 : sb0, jn), nb ? wofs[1] : wofs[0], d0, g0);
#endif
                // (1) resolve step s: energy in the reference's summation order, accept in the log domain
                double En = bcast_d(x_prop, 0);
#pragma unroll
                for (int r = 1; r < R; ++r) En = __dadd_rn(En, bcast_d(x_prop, r));
                const double d = __dsub_rn(E, En);
                const float tt = (float)d - lu_s;
                const bool acc = tt > SPEC_MARGIN;
                const bool unsure = !acc && !(tt < -SPEC_MARGIN);        // guard band, also NaN
                // (2) contributions to step s+1 under both outcomes of step s
#if SPEC_REGCACHE
                const Cand k = evaluate_rc(dw_n, slot_of(nb ? sb1 : sb0, jn), nb ? wofs[1] : wofs[0], d0, g0);
#elif !SPEC_EVAL_FIRST
                const Cand k = evaluate(dw_n, slot_of(nb ? sb1 : sb0, jn), nb ? wofs[1] : wofs[0], d0, g0);
#endif
                rare = rare || unsure || (acc && (int)dw_s >= 0) || k.miss;
                // (3) commit
                commit(acc, s, En, ev_bin[j], ev_cnt[j]);
#if SPEC_REGCACHE
                advance_cache(acc, (int)dw_n < 0);
#endif
                x_prop = acc ? k.xB : k.xA;
                dw_s = dw_n; lu_s = lu_n; d0 = k.d1; g0 = k.g1; own_s = k.own;
#if SPEC_STEP_FENCE
                asm volatile("" ::: "memory");
#endif
            }
            if (__any_sync(SPEC_FULL, rare)) {               // roll the batch back
                E = sv_E; x_prop = sv_x; sg = sv_sg; gm = sv_gm; pg = sv_pg; since_sg = sv_ssg; since_gm = sv_sgm;
                d0 = sv_d0; g0 = sv_g0; acc_total = sv_at; acc_own = sv_ao; dw_s = sv_dw; lu_s = sv_lu; own_s = sv_own;
#pragma unroll
                for (int j = 0; j < BL; ++j) ev_cnt[j] = 0;
                careful = true;
                ++n_rollbacks;
            }
            ++n_batches;
        }

        // ---- careful batch: step by step, every rare case handled exactly -----------------------------
        if (careful) {
#if SPEC_REGCACHE
            cache_stale = true;
#endif
            // generating lane and block of batch position jj (compile-time select chain: no dynamic register index)
            auto g_dw = [&](int jj) { uint32_t v = gen[0][0].dw; _Pragma("unroll") for (int m = 1; m < NB; ++m) v = jj / L == m ? gen[0][m].dw : v; return v; };
            auto g_lu = [&](int jj) { float v = gen[0][0].lu; _Pragma("unroll") for (int m = 1; m < NB; ++m) v = jj / L == m ? gen[0][m].lu : v; return v; };
            auto g_is = [&](int jj) { uint32_t v = gen[0][0].istate; _Pragma("unroll") for (int m = 1; m < NB; ++m) v = jj / L == m ? gen[0][m].istate : v; return v; };
            auto g_u = [&](int jj) { double v = gen[0][0].u; _Pragma("unroll") for (int m = 1; m < NB; ++m) v = jj / L == m ? gen[0][m].u : v; return v; };
#pragma unroll 1
            for (int j = 0; j < BL; ++j) {
                const int s = s0 + j;
                if (D < 3 && j == BL - 1) {
                    cp_async_wait<0>();
                    __syncwarp();
                }
                const bool nb = j + 1 == BL;
                const int jn = nb ? 0 : j + 1;
                const uint32_t dw_n = bcast_u(nb ? gen[1][0].dw : g_dw(jn), jn % L);
                const float lu_n = bcast_f(nb ? gen[1][0].lu : g_lu(jn), jn % L);
                const uint32_t i2n = bcast_u(nb ? gen[1][0].istate : g_is(jn), jn % L);
                const uint32_t i2 = bcast_u(g_is(j), j % L);
                const double u = bcast_d(g_u(j), j % L);
                double En = bcast_d(x_prop, 0);
#pragma unroll
                for (int r = 1; r < R; ++r) En = __dadd_rn(En, bcast_d(x_prop, r));
                const double d = __dsub_rn(E, En);
                const float tt = (float)d - lu_s;
                const bool unsure = !(tt > SPEC_MARGIN) && !(tt < -SPEC_MARGIN);
                bool acc = tt > SPEC_MARGIN;
                if (__any_sync(SPEC_FULL, unsure)) {                      // the specified exact rule, inside the guard band only
                    const bool exact = s < s_end && ((En < E) || (d >= BCP_ACCEPT_CUTOFF && u < exp_det(d)));
                    acc = unsure ? exact : acc;
                }
                const bool state_s = (int)dw_s >= 0;
                uint32_t eb, ec;
                commit(acc, s, En, eb, ec);
                if (ec) atomicAdd(C.param_counts + eb, (unsigned long long)ec);
                const bool accst = acc && state_s;
                if (__any_sync(SPEC_FULL, accst)) {
                    if (accst) {       // adopt the slot as the current state, reload the gamma row
                        const uint32_t slot = slot_of(sb0, j);
                        if (q < R) sts128(cur_rec, lds128(slot + 16 * qr));
                        if (q == 0) Cst = __dadd_rn(lds64(slot + ly.slot_c), logZ);
                        ++acc_state;
                        if ((int)i2 != state) {
                            if (s > burn && lead)
                                atomicAdd(g_state_counts + state, (unsigned long long)(s - since_state));
                            if (s > burn) since_state = s;
                            state = (int)i2;
                        }
                        if (HG) {
                            // the whole row by cp.async: the copies of a lane are all in flight at once (a load /
                            // store loop pays one L2 round trip per 16 bytes: 7 us for a 1.4 KB row at L = 2)
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
                x_prop = evaluate_exact(dw_n, slot_of(nb ? sb1 : sb0, jn), i2n);
                if (s >= burn && s < s_end) {
                    if (TRACE && traced) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
                    if (s == next_snap) {
                        const size_t o = (size_t)snap_idx * n + c;
                        if (lead) {
                            A.traj_E[o] = E;
                            A.traj_state[o] = state;
                            A.traj_accept[o] = acc ? 1 : 0;
                        }
                        if (owner) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = sg;
                        if (HG && glane && live) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = gm;
                        ++snap_idx;
                        const long long ns = (long long)next_snap + A.traj_every;
                        next_snap = ns < s_end ? (int)ns : -1;
                    }
                }
                dw_s = dw_n; lu_s = lu_n; own_s = (dw_n & own_mask) != 0u;
                d0 = (int)((dw_n >> fsh) & 3u) - 1;
                g0 = (int)((dw_n >> 16) & 3u) - 1;
            }
            next_ev = next_snap;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const int sp = s0 + D * BL + m * L + q;
                philox_begin(sp);
#pragma unroll
                for (int r = 0; r < 10; ++r) philox_round(r);
                gnew[m] = philox_finish(sp);
            }
        }

        if (TRACE && !careful && traced) {                    // fast batch: the state did not change
#pragma unroll
            for (int j = 0; j < BL; ++j)
                if (s0 + j >= burn && s0 + j < s_end) A.state_trace[(size_t)(A.s_begin + s0 + j - A.burn) * A.trace_n + c] = state;
        }
        // ---- end of batch: compact the histogram events of the warp into its queue and flush them with
        //      one RED per 32 events; prefetch the state moves of the generated batch, rotate ---------------
        {
            uint32_t evn = 0;
#pragma unroll
            for (int j = 0; j < BL; ++j) {
                const uint32_t m = __ballot_sync(SPEC_FULL, ev_cnt[j] != 0u);
                if (ev_cnt[j] != 0u) {
                    const uint32_t at = evq + 8u * (evn + (uint32_t)__popc(m & lane_lt));
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(at), "r"(ev_bin[j]), "r"(ev_cnt[j]) : "memory");
                }
                evn += (uint32_t)__popc(m);
            }
            __syncwarp();
            for (uint32_t k = threadIdx.x & 31; k < evn; k += 32) {
                uint32_t bin, cnt;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bin), "=r"(cnt) : "r"(evq + 8u * k));
                atomicAdd(C.param_counts + bin, (unsigned long long)cnt);
            }
#pragma unroll
            for (int m = 0; m < NB; ++m) prefetch(gnew[m], slot_of(sb2, m * L + q), gm);
            cp_async_commit();
            wofs[D] = window_ofs();
#pragma unroll
            for (int b = 0; b + 1 < D; ++b)
#pragma unroll
                for (int m = 0; m < NB; ++m) gen[b][m] = gen[b + 1][m];
#pragma unroll
            for (int m = 0; m < NB; ++m) gen[D - 1][m] = gnew[m];
#pragma unroll
            for (int b = 0; b < D; ++b) wofs[b] = wofs[b + 1];
            sb0 = sb1;
        }
    }
    cp_async_wait<0>();

    if ((threadIdx.x & 31) == 0 && n_batches) {
        atomicAdd(C.spec_stats, (unsigned long long)n_batches);
        if (n_rollbacks) atomicAdd(C.spec_stats + 1, (unsigned long long)n_rollbacks);
    }
    // ---- flush run-length counters, persist the chain -------------------------------------
    if (lead) {
        if (s_end > since_state) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - since_state));
        C.state[c] = state;
        C.E[c] = E;
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
}

template <int L, int R, bool HG, bool TRACE>
static cudaError_t launch_spec_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st,
                                   int sms) {
    const SpecLayout ly = spec_layout(SpecCfg<L>::NSLOT, R, HG, T.gh4_stride);
    const long long threads = C.n_chains * L;
    const long long n_warps = (threads + 31) / 32;
    int wpb = (int)((n_warps + sms - 1) / sms);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);      // up to 8 warps per CTA: two waves of 8-lane chains (C2) stay resident per SM
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    size_t smem = 0;
    for (; wpb >= 1; --wpb) {
        smem = (size_t)T.sig4_bytes + (size_t)(wpb * 32 / L) * ly.chain_bytes + (size_t)wpb * 1024 + 128;
        if (smem + 64 <= (size_t)max_optin) break;
    }
    if (wpb < 1) return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(mcmc_spec_kernel<L, R, HG, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int block = wpb * 32;
    const int grid = (int)((threads + block - 1) / block);
    mcmc_spec_kernel<L, R, HG, TRACE><<<grid, block, smem, st>>>(T, C, A);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_spec(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    if (!T.canonical || !T.div_safe || T.R < 1 || T.R > 8 || T.S < 1) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
    const bool tr = A.state_trace != nullptr;
#define BCP_SPEC(l, r)                                                                              \
    case r:                                                                                         \
        return tr ? (hg ? launch_spec_one<l, r, true, true>(T, C, A, st, sms) : launch_spec_one<l, r, false, true>(T, C, A, st, sms)) \
                  : (hg ? launch_spec_one<l, r, true, false>(T, C, A, st, sms) : launch_spec_one<l, r, false, false>(T, C, A, st, sms));
    switch (T.R) {          // L = smallest of 2, 4, 8 that holds one lane per restraint
        BCP_SPEC(2, 1) BCP_SPEC(2, 2) BCP_SPEC(4, 3) BCP_SPEC(4, 4)
        BCP_SPEC(8, 5) BCP_SPEC(8, 6) BCP_SPEC(8, 7) BCP_SPEC(8, 8)
        default: return cudaErrorNotSupported;
    }
#undef BCP_SPEC
}

}  // namespace bcp
