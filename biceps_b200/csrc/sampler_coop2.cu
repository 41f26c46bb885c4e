// K5, cooperative Metropolis kernel, TWO steps per loop iteration.
//
// Like sampler_coop.cu: L lanes per chain, lane q owns restraint q.  With few chains per GPU the
// step time is the length of one warp's (in-order) instruction stream, ~4.7 cycles per instruction.
// This kernel halves the stream per step by resolving steps a = 2k and b = 2k+1 in one iteration:
//   * step b's proposal is known from the counter-based RNG, only its starting configuration depends
//     on whether step a is accepted.  Each lane therefore evaluates THREE restraint terms from
//     independent operands -- t_a, t_b|a rejected, t_b|a accepted -- three independent IEEE divisions
//     that overlap in the fp64 pipe, exchanges them in one shuffle phase and builds the three ordered
//     energy sums; then  acc_a -> (E, E_b candidate) -> acc_b  is a short scalar chain.  Nothing is
//     speculated away: both outcomes of step a are prepared, so every iteration completes two steps.
//   * the gamma restraint keeps the whole sse row of the CURRENT state in shared memory (refilled
//     asynchronously, cp.async, when an accepted move changes the state), so a gamma proposal is one
//     LDS at any index and no sliding-window bookkeeping is needed;
//   * the table words of proposed states (record slice, energy, 7-wide gamma window) are staged two
//     steps ahead into shared memory with cp.async: no registers, no ping-pong code duplication.
// Same RNG stream, same fp64 operation order, bit-identical outputs (tests run every parity case
// through this kernel too).  Canonical layout (tables.cuh), R <= L in {2, 4, 8}.
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

__device__ __forceinline__ int wrap_pm1(int v, int n) {      // v in [-1, n]
    return v < 0 ? v + n : (v >= n ? v - n : v);
}
__device__ __forceinline__ int wrap_pm2(int v, int n) {      // v in [-2, n+1], n >= 2
    return v < 0 ? v + n : (v >= n ? v - n : v);
}
template <int L>
__device__ __forceinline__ double sh_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src, L);
}
template <int L>
__device__ __forceinline__ uint32_t sh_u(uint32_t v, int src) {
    return __shfl_sync(0xffffffffu, v, src, L);
}
__device__ __forceinline__ double u2_bits(uint32_t w2, uint32_t w3) {   // == philox.cuh:u2_from_words
    const uint32_t hi = 0x3FF00000u | (w2 >> 12);
    const uint32_t lo = (w2 << 20) | (w3 >> 12);
    return __dadd_rn(__dadd_rn(__hiloint2double((int)hi, (int)lo), -1.0), 1.1102230246251565e-16);
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// per-lane staging slot of one prefetched state move (shared memory)
struct alignas(16) Stage {
    double2 rec;     // {sse_q, ref_q} of the proposed state
    double C;        // energy of the proposed state (lambda row of the chain)
    double pw[7];    // gamma lane: sse[i'][jgp-3 .. jgp+3]
};

template <int L, int R, bool HG>
__global__ void __launch_bounds__(128)
mcmc_coop2_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                  const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);

    const int G = HG ? T.r[R - 1].n_gamma : 1;
    const int Gp = HG ? T.gamma_row_stride : 0;
    const int CPB = blockDim.x / L;                                  // chains per CTA
    Stage* s_stage = reinterpret_cast<Stage*>(smem_raw + T.sig_bytes);            // [4][blockDim]
    double* s_rows = reinterpret_cast<double*>(s_stage + 4 * blockDim.x);          // [CPB][Gp]

    const long long n = C.n_chains;
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(threadIdx.x & (L - 1));
    const int cl = threadIdx.x / L;                                  // chain within the CTA
    long long c = gthread / L;
    const bool live = c < n;
    if (!live) c = n - 1;
    const int qr = q < R ? q : 0;
    const bool owner = live && q < R;
    const bool lead = live && q == 0;
    const bool glane = HG && q == R - 1;
    (void)CPB;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(T.rec) + qr;
    const double* __restrict__ halo3 = HG ? T.gsse_halo3 : nullptr;          // rows of G + 6
    const double* __restrict__ grows = HG ? T.gsse_rows : nullptr;           // rows of Gp, 16-byte aligned
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    unsigned long long* g_sigma_hist = C.param_counts + (size_t)grp * T.HB + T.r[qr].hist_sigma;
    unsigned long long* g_gamma_hist = C.param_counts + (size_t)grp * T.HB + (HG ? T.r[R - 1].hist_gamma : 0);
    const uint32_t S = (uint32_t)T.S;
    const uint32_t thresh = T.nuis_thresh;
    const int s_end = (int)(A.s_end - A.s_begin);
    const long long burn_l = A.burn - A.s_begin;
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
    const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
    const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);

    const int nsig = T.r[qr].n_sigma;
    const double* __restrict__ s_a = s_sig + T.r[qr].sig_off;
    const double* __restrict__ s_d = s_a + nsig;
    const double cnorm = T.r[qr].cnorm;
    double* my_row = HG ? s_rows + (size_t)cl * Gp : nullptr;
    Stage* my_stage = s_stage + threadIdx.x;                         // + slot * blockDim.x

    // ---- chain state -----------------------------------------------------------------------
    int state = C.state[c];
    double E = C.E[c];
    int sg = C.isg[(size_t)qr * n + c];
    int jg = HG ? C.igm[(size_t)(R - 1) * n + c] : 0;                // replicated in every lane
    double csse, cref;
    {
        const double2 v = rec2[(size_t)state * R];
        csse = v.x;
        cref = v.y;
    }
    double Cst = __dadd_rn(energy[state], logZ);
    int since_sg = 0, since_gm = 0, since_state = 0;
    unsigned int acc_own = 0, acc_state = 0, acc_total = 0;
    bool row_pending = false;
    int drift = 0;          // gamma index now minus the centre of this iteration's staged windows (|drift| <= 2)

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }

    // cooperative asynchronous copy of the gamma row of `st` into this chain's shared row
    auto load_row = [&](int st) {
        if (HG) {
            const char* src = reinterpret_cast<const char*>(grows + (size_t)st * Gp);
            char* dst = reinterpret_cast<char*>(my_row);
            for (int k = q * 16; k < Gp * 8; k += L * 16) cp_async16(dst + k, src + k);
        }
    };
    // stage the table words a state move to (w1 * S) >> 32 needs, centred on the current gamma index
    auto prefetch = [&](int slot, uint32_t w1) {
        const uint32_t i2 = __umulhi(w1, S);
        Stage* st = my_stage + slot * blockDim.x;
        cp_async16(&st->rec, rec2 + (size_t)i2 * R);
        cp_async8(&st->C, energy + i2);
        if (HG && glane) {
            const double* row = halo3 + (size_t)i2 * (G + 6) + jg;     // elements jg-3 .. jg+3
#pragma unroll
            for (int k = 0; k < 7; ++k) cp_async8(&st->pw[k], row + k);
        }
    };
    auto block_of = [&](int s) {
        const unsigned long long ctr = ctr0 + (unsigned long long)s;
        return philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), ch_lo, ch_hi, seed_lo, seed_hi);
    };
    auto bcast = [&](const Philox4& b, int j) {
        Philox4 r;
        r.w0 = sh_u<L>(b.w0, j); r.w1 = sh_u<L>(b.w1, j);
        r.w2 = sh_u<L>(b.w2, j); r.w3 = sh_u<L>(b.w3, j);
        return r;
    };

    // bookkeeping of one resolved step (lane-local; `acc` etc. are replicated in the group)
    auto commit = [&](const int s, const bool acc, const bool nuis, const int hi, const bool mine, const int dsg,
                      const int dgm, const double En, const double Cn, const double2 prec) {
        acc_total += acc ? 1u : 0u;
        acc_own += (acc && mine) ? 1u : 0u;
        if (acc) {
            E = En;
            if (nuis) {
                if (mine) {
                    const int sgn = wrap_pm1(sg + dsg, nsig);
                    if (sgn != sg) {
                        const int lo = since_sg > burn ? since_sg : burn;
                        if (owner && s > lo) atomicAdd(g_sigma_hist + sg, (unsigned long long)(s - lo));
                        since_sg = s;
                        sg = sgn;
                    }
                }
                if (HG && dgm != 0) {                        // dgm is replicated: every lane tracks jg
                    if (glane) {
                        const int lo = since_gm > burn ? since_gm : burn;
                        if (live && s > lo) atomicAdd(g_gamma_hist + jg, (unsigned long long)(s - lo));
                        since_gm = s;
                    }
                    jg = wrap_pm1(jg + dgm, G);
                }
            } else {
                ++acc_state;
                csse = prec.x;
                cref = prec.y;
                Cst = Cn;
                if (hi != state) {
                    const int lo = since_state > burn ? since_state : burn;
                    if (lead && s > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s - lo));
                    since_state = s;
                    state = hi;
                }
            }
        }
        if (s >= burn) {
            if (A.state_trace && lead) A.state_trace[(size_t)(A.s_begin + s - A.burn) * n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                if (lead) {
                    A.traj_E[o] = E;
                    A.traj_state[o] = state;
                    A.traj_accept[o] = acc ? 1 : 0;
                }
                if (owner) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = sg;
                if (HG && glane && live) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = jg;
                ++snap_idx;
                const long long ns = (long long)next_snap + A.traj_every;
                next_snap = ns < s_end ? (int)ns : -1;
            }
        }
    };

    // ---- prologue: gamma row of the starting state, RNG and staging for steps 0 and 1 ---------
    Philox4 wbatch = block_of(q);                       // lane j holds the block of step s0 + j
    Philox4 wa = bcast(wbatch, 0), wb = bcast(wbatch, 1 % L);
    load_row(state);
    if (!(wa.w0 < thresh)) prefetch(0, wa.w1);
    if (!(wb.w0 < thresh)) prefetch(1, wb.w1);
    cp_commit();
    cp_wait<0>();
    __syncwarp();

    for (int s0 = 0; s0 < s_end; s0 += L) {
        const Philox4 wnext = block_of(s0 + L + q);
#pragma unroll 1
        for (int j = 0; j < L; j += 2) {
            const int sa = s0 + j, sb = sa + 1;
            if (sa >= s_end) break;                                  // uniform
            const bool valid_b = sb < s_end;
            const int slot_a = sa & 3, slot_b = sb & 3;

            // ---- RNG of the NEXT iteration's steps, staged immediately (two steps of lead time) ----
            const Philox4 na = (j + 2 < L) ? bcast(wbatch, j + 2) : bcast(wnext, 0);
            const Philox4 nb = (j + 3 < L) ? bcast(wbatch, j + 3) : bcast(wnext, (j + 3) - L);
            // every copy staged for THIS iteration was committed one iteration ago: all but the newest
            // group (the row refill, possibly empty) must have landed
            cp_wait<1>();
            __syncwarp();
            if (sa + 2 < s_end && !(na.w0 < thresh)) prefetch(slot_a ^ 2, na.w1);
            if (sb + 2 < s_end && !(nb.w0 < thresh)) prefetch(slot_b ^ 2, nb.w1);
            cp_commit();

            // ---- decode both proposals ---------------------------------------------------------
            const bool nuis_a = wa.w0 < thresh, nuis_b = wb.w0 < thresh;
            const uint64_t pra = (uint64_t)wa.w1 * (nuis_a ? (uint32_t)R : S);
            const uint64_t prb = (uint64_t)wb.w1 * (nuis_b ? (uint32_t)R : S);
            const int hia = (int)(pra >> 32), hib = (int)(prb >> 32);
            const bool mine_a = nuis_a && hia == q, mine_b = nuis_b && hib == q;
            const uint64_t za1 = (uint64_t)(uint32_t)pra * 3u, zb1 = (uint64_t)(uint32_t)prb * 3u;
            const int dsg_a = (int)(za1 >> 32) - 1, dsg_b = (int)(zb1 >> 32) - 1;
            int dgm_a = 0, dgm_b = 0;                     // replicated: nonzero iff the gamma restraint is picked
            if (HG) {
                const uint64_t za2 = (uint64_t)(uint32_t)za1 * 3u, zb2 = (uint64_t)(uint32_t)zb1 * 3u;
                dgm_a = (nuis_a && hia == R - 1) ? (int)(za2 >> 32) - 1 : 0;
                dgm_b = (nuis_b && hib == R - 1) ? (int)(zb2 >> 32) - 1 : 0;
            }
            // the gamma row is being refilled (state changed last iteration) and is needed now: wait for it
            if (HG) {
                const bool need_row = row_pending && ((nuis_a && hia == R - 1) || (nuis_b && hib == R - 1));
                if (__any_sync(0xffffffffu, need_row)) { cp_wait<0>(); __syncwarp(); }
            }

            // ---- operands of the three terms ---------------------------------------------------
            const Stage* Pa = my_stage + slot_a * blockDim.x;
            const Stage* Pb = my_stage + slot_b * blockDim.x;
            const double2 ra = Pa->rec, rb = Pb->rec;
            const int sA = mine_a ? wrap_pm1(sg + dsg_a, nsig) : sg;            // sigma index after a accepted
            const int sBr = mine_b ? wrap_pm1(sg + dsg_b, nsig) : sg;           // step b, a rejected
            const int sBa = mine_b ? wrap_pm1(sA + dsg_b, nsig) : sA;           // step b, a accepted
            double num_a = nuis_a ? csse : ra.x;
            double num_br = nuis_b ? csse : rb.x;
            double num_ba = nuis_b ? (nuis_a ? csse : ra.x) : rb.x;
            if (HG && glane) {
                const int gA = wrap_pm1(jg + dgm_a, G);
                const int ctr = 3 + drift;               // window slot of the current gamma index
                num_a = nuis_a ? my_row[gA] : Pa->pw[ctr];
                num_br = nuis_b ? my_row[wrap_pm1(jg + dgm_b, G)] : Pb->pw[ctr];
                num_ba = nuis_b ? (nuis_a ? my_row[wrap_pm2(jg + dgm_a + dgm_b, G)] : Pa->pw[ctr + dgm_b])
                                : Pb->pw[ctr + dgm_a];
            }
            const double ref_a = nuis_a ? cref : ra.y;
            const double ref_br = nuis_b ? cref : rb.y;
            const double ref_ba = nuis_b ? ref_a : rb.y;
            const double C_a = nuis_a ? Cst : __dadd_rn(Pa->C, logZ);
            const double C_bs = __dadd_rn(Pb->C, logZ);
            const double C_br = nuis_b ? Cst : C_bs;
            const double C_ba = nuis_b ? C_a : C_bs;

            // ---- three independent terms, reference fp64 order ----------------------------------
            double t_a = __dadd_rn(s_a[sA], __ddiv_rn(num_a, s_d[sA]));
            double t_br = __dadd_rn(s_a[sBr], __ddiv_rn(num_br, s_d[sBr]));
            double t_ba = __dadd_rn(s_a[sBa], __ddiv_rn(num_ba, s_d[sBa]));
            t_a = __dsub_rn(__dadd_rn(t_a, cnorm), ref_a);
            t_br = __dsub_rn(__dadd_rn(t_br, cnorm), ref_br);
            t_ba = __dsub_rn(__dadd_rn(t_ba, cnorm), ref_ba);
            double En_a = C_a, En_br = C_br, En_ba = C_ba;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                En_a = __dadd_rn(En_a, sh_d<L>(t_a, k));
                En_br = __dadd_rn(En_br, sh_d<L>(t_br, k));
                En_ba = __dadd_rn(En_ba, sh_d<L>(t_ba, k));
            }

            // ---- resolve a, then b ---------------------------------------------------------------
            const int state0 = state;
            int new_drift = 0;
            const bool acc_a = (En_a < E) || metropolis_accept(__dsub_rn(E, En_a), u2_bits(wa.w2, wa.w3));
            commit(sa, acc_a, nuis_a, hia, mine_a, dsg_a, dgm_a, En_a, C_a, ra);
            new_drift += acc_a ? dgm_a : 0;
            if (valid_b) {
                const double En_b = acc_a ? En_ba : En_br;
                const double C_b = acc_a ? C_ba : C_br;
                const bool acc_b = (En_b < E) || metropolis_accept(__dsub_rn(E, En_b), u2_bits(wb.w2, wb.w3));
                commit(sb, acc_b, nuis_b, hib, mine_b, dsg_b, dgm_b, En_b, C_b, rb);
                new_drift += acc_b ? dgm_b : 0;
            }
            drift = new_drift;        // the next iteration's windows were centred on this iteration's starting gamma

            // ---- keep the shared gamma row in step with the chain's state -------------------------
            if (HG) {
                const bool changed = state != state0;
                if (__any_sync(0xffffffffu, changed && row_pending)) { cp_wait<0>(); __syncwarp(); }
                if (changed) load_row(state);
                row_pending = changed;
            }
            cp_commit();

            wa = na;
            wb = nb;
        }
        wbatch = wnext;
    }
    cp_wait<0>();

    // ---- flush run-length counters, persist the chain -------------------------------------
    if (lead) {
        const int lo = since_state > burn ? since_state : burn;
        if (s_end > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - lo));
        C.state[c] = state;
        C.E[c] = E;
        C.accepted[c] += acc_total;
        C.sep[(size_t)R * n + c] += acc_state;
    }
    if (owner) {
        const int lo = since_sg > burn ? since_sg : burn;
        if (s_end > lo) atomicAdd(g_sigma_hist + sg, (unsigned long long)(s_end - lo));
        C.isg[(size_t)q * n + c] = sg;
        C.sep[(size_t)q * n + c] += acc_own;
    }
    if (HG && glane && live) {
        const int lo = since_gm > burn ? since_gm : burn;
        if (s_end > lo) atomicAdd(g_gamma_hist + jg, (unsigned long long)(s_end - lo));
        C.igm[(size_t)(R - 1) * n + c] = jg;
    }
}

template <int L, int R, bool HG>
static cudaError_t launch_coop2_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st,
                                    int sms) {
    const long long threads = C.n_chains * L;
    const long long n_warps = (threads + 31) / 32;
    int wpb = (int)((n_warps + sms - 1) / sms);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const int block = wpb * 32;
    const int grid = (int)((threads + block - 1) / block);
    const size_t smem = (size_t)T.sig_bytes + (size_t)4 * block * sizeof(Stage) +
                        (HG ? (size_t)(block / L) * T.gamma_row_stride * sizeof(double) : 0);
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(mcmc_coop2_kernel<L, R, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mcmc_coop2_kernel<L, R, HG><<<grid, block, smem, st>>>(T, C, A);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_coop2(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    if (!T.canonical || T.R < 1 || T.R > 8 || (T.sig_bytes & 15)) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
#define BCP_COOP2(l, r)                                                                             \
    case r:                                                                                         \
        return hg ? launch_coop2_one<l, r, true>(T, C, A, st, sms) : launch_coop2_one<l, r, false>(T, C, A, st, sms);
    switch (T.R) {
        BCP_COOP2(2, 1) BCP_COOP2(2, 2) BCP_COOP2(4, 3) BCP_COOP2(4, 4)
        BCP_COOP2(8, 5) BCP_COOP2(8, 6) BCP_COOP2(8, 7) BCP_COOP2(8, 8)
        default: return cudaErrorNotSupported;
    }
#undef BCP_COOP2
}

}  // namespace bcp
