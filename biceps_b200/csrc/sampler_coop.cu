// K5, cooperative Metropolis kernel: L lanes per chain (L = 2, 4 or 8 >= R), one lane per restraint.
//
// A single chain is a strictly serial dependency chain, so with few chains per GPU (the C3
// configuration runs 4096) the step time is the length of ONE warp's instruction stream.  The
// thread-per-chain kernels re-evaluate all R restraint terms in every lane (R IEEE divisions,
// R-way select chains, a whole Philox block per lane and step): ~515 warp instructions per step.
// Here the independent pieces of a step are spread over the lanes of the chain's group:
//   * lane q owns restraint q: its sigma index, the state-dependent words (sse_q, ref_q or the
//     gamma window) and its histogram bookkeeping; it evaluates ONE term (one division);
//   * the R terms are exchanged with shuffles and summed by every lane in the reference's
//     left-to-right order (bit-identical energies), the accept decision is replicated;
//   * lane j computes the Philox block of step s0 + j for a batch of L steps (1/L block per step);
//   * lane q prefetches its own 16-byte slice of the next proposed state's record.
// Same RNG stream, same fp64 operation order, same outputs as the other kernels (oracle parity
// is bit for bit); canonical restraint layout only (tables.cuh).
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

__device__ __forceinline__ int cwrap1(int v, int n) {      // v in [-1, n]
    return v < 0 ? v + n : (v >= n ? v - n : v);
}
__device__ __forceinline__ double csel3(int d, double m, double z, double p) {   // d in {-1,0,1}
    return d < 0 ? m : (d > 0 ? p : z);
}
template <int L>
__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src, L);
}
template <int L>
__device__ __forceinline__ uint32_t shfl_u(uint32_t v, int src) {
    return __shfl_sync(0xffffffffu, v, src, L);
}

struct CoopPrefetch {
    double C;
    double2 rec;
    double W[5];
};

// Metropolis uniform (2k+1) * 2^-53 from the 52-bit integer k = (w2 << 20) | (w3 >> 12), built with
// bit operations and two exact fp64 adds: 1.k - 1.0 = k * 2^-52, + 2^-53  (== philox.cuh:u2_from_words)
__device__ __forceinline__ double u2_fast(uint32_t w2, uint32_t w3) {
    const uint32_t hi = 0x3FF00000u | (w2 >> 12);
    const uint32_t lo = (w2 << 20) | (w3 >> 12);
    const double one_k = __hiloint2double((int)hi, (int)lo);
    return __dadd_rn(__dadd_rn(one_k, -1.0), 1.1102230246251565e-16);
}

template <int L, int R, bool HG>
__global__ void __launch_bounds__(128)
mcmc_coop_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                 const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);

    const long long n = C.n_chains;
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(threadIdx.x & (L - 1));             // lane within the chain's group
    long long c = gthread / L;
    const bool live = c < n;                                // dead groups shadow the last chain, no stores
    if (!live) c = n - 1;
    const int qr = q < R ? q : 0;                           // restraint this lane evaluates (idle lanes: 0)
    const bool owner = live && q < R;                       // lane owns restraint q
    const bool lead = live && q == 0;
    const bool glane = HG && q == R - 1;                    // lane of the gamma restraint

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(T.rec) + qr;    // this lane's slice
    const double* __restrict__ gsse = HG ? T.gsse_halo : nullptr;
    const int G = HG ? T.r[R - 1].n_gamma : 1;
    const int GH = G + 4;
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

    // this lane's restraint: sigma vectors in shared memory, normalisation constant
    const int nsig = T.r[qr].n_sigma;
    const double* __restrict__ s_a = s_sig + T.r[qr].sig_off;      // Ndof ln(sigma)
    const double* __restrict__ s_d = s_a + nsig;                   // 2 sigma^2
    const double cnorm = T.r[qr].cnorm;

    // ---- chain state (state, E, Cst replicated in every lane of the group) ----------------
    int state = C.state[c];
    double E = C.E[c];
    int sg = C.isg[(size_t)qr * n + c];
    int jg = 0;                                              // gamma index (gamma lane only)
    double csse, cref, W0 = 0.0, W1 = 0.0, W2 = 0.0;
    {
        const double2 v = rec2[(size_t)state * R];
        csse = v.x;
        cref = v.y;
    }
    if (HG) {
        jg = C.igm[(size_t)(R - 1) * n + c];
        const double* row = gsse + (size_t)state * GH + jg + 1;
        W0 = row[0]; W1 = row[1]; W2 = row[2];
    }
    double Cst = __dadd_rn(energy[state], logZ);
    int since_sg = 0, since_gm = 0, since_state = 0, pshift = 0;
    unsigned int acc_own = 0, acc_state = 0, acc_total = 0;

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }

    auto prefetch = [&](CoopPrefetch& p, uint32_t n1) {
        const uint32_t i2 = __umulhi(n1, S);
        p.C = __ldg(energy + i2);
        p.rec = __ldg(rec2 + (size_t)i2 * R);
        if (HG && glane) {
            const double* row = gsse + (size_t)i2 * GH + jg;            // halo: [jg-2 .. jg+2]
#pragma unroll
            for (int k = 0; k < 5; ++k) p.W[k] = __ldg(row + k);
        }
    };

    // one step; w = this step's Philox block, nw = the next step's (prefetch decision)
    auto step = [&](const int s, const Philox4& w, const Philox4& nw, const bool has_next, const CoopPrefetch& use,
                    CoopPrefetch& fill) {
        const bool nuis = w.w0 < thresh;
        const uint64_t pr = (uint64_t)w.w1 * (nuis ? (uint32_t)R : S);
        const int hi = (int)(pr >> 32);
        const bool mine = nuis && hi == q;
        const uint64_t z1 = (uint64_t)(uint32_t)pr * 3u;
        const int dsg = (int)(z1 >> 32) - 1;
        int dgm = 0;
        if (HG) {
            const uint64_t z2 = (uint64_t)(uint32_t)z1 * 3u;
            dgm = (mine && glane) ? (int)(z2 >> 32) - 1 : 0;
        }
        const int sgn = mine ? cwrap1(sg + dsg, nsig) : sg;
        const double av = s_a[sgn];
        const double dv = s_d[sgn];
        double num = nuis ? csse : use.rec.x;
        const double ref = nuis ? cref : use.rec.y;
        const double Cn = nuis ? Cst : __dadd_rn(use.C, logZ);
        const double keep_x = use.rec.x, keep_y = use.rec.y;
        double nW0 = 0.0, nW2 = 0.0;
        if (HG) {
            // the window prefetched for a state move is centred on the gamma index of one step ago;
            // only if that step moved gamma (pshift != 0, rare) does the centre have to be re-selected
            double wc = use.W[2];
            nW0 = use.W[1];
            nW2 = use.W[3];
            if (__any_sync(0xffffffffu, pshift != 0)) {
                wc = csel3(pshift, use.W[1], use.W[2], use.W[3]);
                nW0 = csel3(pshift, use.W[0], use.W[1], use.W[2]);
                nW2 = csel3(pshift, use.W[2], use.W[3], use.W[4]);
            }
            if (glane) num = nuis ? csel3(dgm, W0, W1, W2) : wc;
        }

        if (has_next && !(nw.w0 < thresh)) prefetch(fill, nw.w1);

        double t = __dadd_rn(av, __ddiv_rn(num, dv));
        t = __dadd_rn(t, cnorm);
        t = __dsub_rn(t, ref);
        // every lane sums the R terms in the reference's order
        double En = Cn;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const double tk = shfl_d<L>(t, k);
            if (k < R) En = __dadd_rn(En, tk);
        }
        const bool acc = (En < E) || metropolis_accept(__dsub_rn(E, En), u2_fast(w.w2, w.w3));

        E = acc ? En : E;
        acc_total += acc ? 1u : 0u;
        acc_own += (acc && mine) ? 1u : 0u;
        pshift = 0;
        if (acc) {
            if (nuis) {
                if (mine) {
                    if (sgn != sg) {
                        const int lo = since_sg > burn ? since_sg : burn;
                        if (owner && s > lo) atomicAdd(g_sigma_hist + sg, (unsigned long long)(s - lo));
                        since_sg = s;
                        sg = sgn;
                    }
                    if (HG && dgm != 0) {
                        const int lo = since_gm > burn ? since_gm : burn;
                        if (live && s > lo) atomicAdd(g_gamma_hist + jg, (unsigned long long)(s - lo));
                        const int gmn = cwrap1(jg + dgm, G);
                        const double edge = __ldg(gsse + (size_t)state * GH + (gmn + 2 + dgm));
                        if (dgm > 0) { W0 = W1; W1 = W2; W2 = edge; }
                        else { W2 = W1; W1 = W0; W0 = edge; }
                        since_gm = s;
                        jg = gmn;
                        pshift = dgm;
                    }
                }
            } else {
                ++acc_state;
                csse = keep_x;
                cref = keep_y;
                Cst = Cn;
                if (HG && glane) { W0 = nW0; W1 = num; W2 = nW2; }
                if (hi != state) {
                    const int lo = since_state > burn ? since_state : burn;
                    if (lead && s > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s - lo));
                    since_state = s;
                    state = hi;
                }
            }
        }

        if (s >= burn) {
            if (A.state_trace && lead && c < A.trace_n) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
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

    // ---- batches of L steps: lane j generates the Philox block of step s0 + j; two steps per
    //      loop iteration so the prefetch buffers ping-pong without register copies --------------
    auto block_of = [&](int s) {
        const unsigned long long ctr = ctr0 + (unsigned long long)s;
        return philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), ch_lo, ch_hi, seed_lo, seed_hi);
    };
    auto bcast = [&](const Philox4& b, int j) {
        Philox4 r;
        r.w0 = shfl_u<L>(b.w0, j); r.w1 = shfl_u<L>(b.w1, j);
        r.w2 = shfl_u<L>(b.w2, j); r.w3 = shfl_u<L>(b.w3, j);
        return r;
    };
    Philox4 wb = block_of(q);
    CoopPrefetch pa, pb;
    pa.C = 0.0; pa.rec = make_double2(0.0, 0.0); pb.C = 0.0; pb.rec = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 5; ++k) { pa.W[k] = 0.0; pb.W[k] = 0.0; }
    Philox4 wa = bcast(wb, 0);                                   // block of step 0
    if (!(wa.w0 < thresh)) prefetch(pa, wa.w1);
    for (int s0 = 0; s0 < s_end; s0 += L) {
        const Philox4 wnext = block_of(s0 + L + q);
        for (int j = 0; j < L; j += 2) {                         // L is even
            const int s = s0 + j;
            if (s >= s_end) break;                               // uniform
            const Philox4 w1 = bcast(wb, j + 1);
            step(s, wa, w1, s + 1 < s_end, pa, pb);
            if (s + 1 >= s_end) break;
            wa = (j + 2 < L) ? bcast(wb, j + 2) : bcast(wnext, 0);
            step(s + 1, w1, wa, s + 2 < s_end, pb, pa);
        }
        wb = wnext;
    }

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
static cudaError_t launch_coop_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st,
                                   int sms) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_coop_kernel<L, R, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.sig_bytes);
    if (e != cudaSuccess) return e;
    const long long threads = C.n_chains * L;
    const long long n_warps = (threads + 31) / 32;
    int wpb = (int)((n_warps + sms - 1) / sms);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const int block = wpb * 32;
    const int grid = (int)((threads + block - 1) / block);
    mcmc_coop_kernel<L, R, HG><<<grid, block, T.sig_bytes, st>>>(T, C, A);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_coop(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms) {
    if (!T.canonical || T.R < 1 || T.R > 8) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
#define BCP_COOP(l, r)                                                                              \
    case r:                                                                                         \
        return hg ? launch_coop_one<l, r, true>(T, C, A, st, sms) : launch_coop_one<l, r, false>(T, C, A, st, sms);
    switch (T.R) {          // L = smallest of 2, 4, 8 that holds one lane per restraint
        BCP_COOP(2, 1) BCP_COOP(2, 2) BCP_COOP(4, 3) BCP_COOP(4, 4)
        BCP_COOP(8, 5) BCP_COOP(8, 6) BCP_COOP(8, 7) BCP_COOP(8, 8)
        default: return cudaErrorNotSupported;
    }
#undef BCP_COOP
}

}  // namespace bcp
