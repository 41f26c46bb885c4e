// K5, specialised Metropolis kernel for the canonical restraint layout (every restraint SCALAR
// except, optionally, the LAST one which carries a gamma grid -- the reference's file-extension
// order cs_H, cs_Ca, cs_N, J, noe always yields this layout).
//
// Same algorithm, RNG stream and fp64 operation order as the generic kernel in sampler.cu (and
// as oracle/sampler_oracle.py): every output is bit-identical.  What changes is the execution
// shape, which is what the generic kernel loses most of its time on (ncu: 839 warp
// instructions per warp-step at 39 % lane utilisation, one L2 round trip per divergent path):
//   * uniform control flow: every lane re-evaluates all R restraint terms each step from
//     register-resident operands (R independent IEEE divisions -> ILP, no divergence between
//     "state move" and "nuisance move on restraint q" lanes).  Re-evaluating an unchanged term
//     from identical operands reproduces its bits, so no per-restraint cache is needed;
//   * the state-dependent words of the CURRENT state live in registers (sse_q, ref_q, a 3-wide
//     gamma window), so a nuisance move touches no global memory at all;
//   * the proposed state of step s+1 is known from the counter-based RNG before step s is
//     resolved: its packed record (rec[i'], one or two sectors), energy and a 5-wide gamma
//     window are loaded one step ahead, off the critical path;
//   * Philox for step s+1 is issued alongside the divisions of step s (software pipelining).
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

__device__ __forceinline__ int wrap1(int v, int n) {      // v in [-1, n]
    return v < 0 ? v + n : (v >= n ? v - n : v);
}
__device__ __forceinline__ int inc_wrap(int v, int n) {   // v in [0, n)
    return v + 1 == n ? 0 : v + 1;
}
__device__ __forceinline__ int wrap_any(int v, int n) {   // any v, small |v - [0,n)| distance
    while (v < 0) v += n;
    while (v >= n) v -= n;
    return v;
}
__device__ __forceinline__ double sel3(int d, double m, double z, double p) {   // d in {-1,0,1}
    return d < 0 ? m : (d > 0 ? p : z);
}

template <int R, bool HG>
__global__ void __launch_bounds__(128)
mcmc_fast_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                 const __grid_constant__ LaunchArgs A) {
    constexpr int RS = HG ? R - 1 : R;      // scalar restraints
    constexpr int QG = R - 1;               // index of the gamma restraint (if HG)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);

    const long long n = C.n_chains;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(T.rec);
    const double* __restrict__ gsse = HG ? T.r[QG].sse : nullptr;
    const int G = HG ? T.r[QG].n_gamma : 1;
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
    const int S = T.S;
    const uint32_t thresh = T.nuis_thresh;
    const int s_begin = 0, s_end = (int)(A.s_end - A.s_begin);      // launch-local step counter
    const long long burn_l = A.burn - A.s_begin;                     // may be negative
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;

    // ---- chain state -------------------------------------------------------------------
    int state = C.state[c];
    double E = C.E[c];
    int isg[R];
    int igm = 0;
    double cA[R], cD[R];
    double csse[R], cref[R];            // [QG] entries of csse unused when HG (window instead)
    double W0 = 0.0, W1 = 0.0, W2 = 0.0;
    int since_sg[R];
    int since_gm = s_begin, since_state = s_begin;
    unsigned int acc_r[R];
    unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        isg[q] = C.isg[(size_t)q * n + c];
        cA[q] = s_sig[T.r[q].sig_off + isg[q]];
        cD[q] = s_sig[T.r[q].sig_off + T.r[q].n_sigma + isg[q]];
        const double2 v = rec2[(size_t)state * R + q];
        csse[q] = v.x;
        cref[q] = v.y;
        since_sg[q] = s_begin;
        acc_r[q] = 0;
    }
    if (HG) {
        igm = C.igm[(size_t)QG * n + c];
        const double* row = gsse + (size_t)state * G;
        W0 = row[wrap1(igm - 1, G)];
        W1 = row[igm];
        W2 = row[wrap1(igm + 1, G)];
    }
    double Cst = __dadd_rn(energy[state], logZ);

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }

    // ---- software pipeline prologue: RNG + state-move prefetch for the first step ------
    Philox4 w = step_block(C.seed, chain_id, ctr0);
    double pC = 0.0;
    double2 prec[R];
    double pW[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < R; ++q) prec[q] = make_double2(0.0, 0.0);
    int pshift = 0;
    {
        const bool nuis0 = w.w0 < thresh;
        if (!nuis0) {
            const int i2 = (int)(((uint64_t)w.w1 * (uint32_t)S) >> 32);
            pC = __ldg(energy + i2);
#pragma unroll
            for (int q = 0; q < R; ++q) prec[q] = __ldg(rec2 + (size_t)i2 * R + q);
            if (HG) {
                const double* row = gsse + (size_t)i2 * G;
                int v = wrap_any(igm - 2, G);
#pragma unroll
                for (int k = 0; k < 5; ++k) { pW[k] = __ldg(row + v); v = inc_wrap(v, G); }
            }
        }
    }

    for (int s = s_begin; s < s_end; ++s) {
        // ---- decode the proposal of step s -------------------------------------------
        const bool nuis = w.w0 < thresh;
        const uint64_t pr = (uint64_t)w.w1 * (uint32_t)(nuis ? R : S);
        const int hi = (int)(pr >> 32);
        uint32_t f = (uint32_t)pr;
        const int pick = nuis ? hi : -1;
        const int nstate = nuis ? state : hi;
        uint64_t z = (uint64_t)f * 3u;
        const int dsg = (int)(z >> 32) - 1;
        f = (uint32_t)z;
        z = (uint64_t)f * 3u;
        const bool pg = HG && pick == QG;
        const int dgm = pg ? (int)(z >> 32) - 1 : 0;
        const double u2 = u2_from_words(w.w2, w.w3);

        int sg_old = 0, nsig = 1, off = 0;
#pragma unroll
        for (int q = 0; q < R; ++q)
            if (q == pick) { sg_old = isg[q]; nsig = T.r[q].n_sigma; off = T.r[q].sig_off; }
        const int sgn = nuis ? wrap1(sg_old + dsg, nsig) : 0;
        const double An = s_sig[off + sgn];
        const double Dn = s_sig[off + nsig + sgn];
        const int gmn = HG ? wrap1(igm + dgm, G) : 0;

        // ---- operands of the R terms (registers only) --------------------------------
        double num[R], ref[R];
#pragma unroll
        for (int q = 0; q < RS; ++q) {
            num[q] = nuis ? csse[q] : prec[q].x;
            ref[q] = nuis ? cref[q] : prec[q].y;
        }
        if (HG) {
            const double wsel = sel3(dgm, W0, W1, W2);
            const double psel = sel3(pshift, pW[1], pW[2], pW[3]);
            num[QG] = nuis ? wsel : psel;
            ref[QG] = nuis ? cref[QG] : prec[QG].y;
        }
        const double Cn = nuis ? Cst : __dadd_rn(pC, logZ);
        // state-move words that survive an acceptance (consumed before the next prefetch)
        double nW0 = 0.0, nW2 = 0.0;
        if (HG) {
            nW0 = sel3(pshift, pW[0], pW[1], pW[2]);
            nW2 = sel3(pshift, pW[2], pW[3], pW[4]);
        }
        double2 keep[R];
#pragma unroll
        for (int q = 0; q < R; ++q) keep[q] = prec[q];

        // ---- next step: RNG and, if it is a state move, its table words ---------------
        const Philox4 wn = step_block(C.seed, chain_id, ctr0 + (unsigned long long)(s + 1));
        if (!(wn.w0 < thresh) && s + 1 < s_end) {
            const int i2 = (int)(((uint64_t)wn.w1 * (uint32_t)S) >> 32);
            pC = __ldg(energy + i2);
#pragma unroll
            for (int q = 0; q < R; ++q) prec[q] = __ldg(rec2 + (size_t)i2 * R + q);
            if (HG) {
                const double* row = gsse + (size_t)i2 * G;
                int v = wrap_any(igm - 2, G);
#pragma unroll
                for (int k = 0; k < 5; ++k) { pW[k] = __ldg(row + v); v = inc_wrap(v, G); }
            }
        }

        // ---- energy of the proposal, reference fp64 order ----------------------------
        double En = Cn;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const double a = (q == pick) ? An : cA[q];
            const double d = (q == pick) ? Dn : cD[q];
            double t = __dadd_rn(a, __ddiv_rn(num[q], d));
            t = __dadd_rn(t, T.r[q].cnorm);
            t = __dsub_rn(t, ref[q]);
            En = __dadd_rn(En, t);
        }

        bool acc;
        if (En < E) acc = true;
        else acc = metropolis_accept(__dsub_rn(E, En), u2);

        // ---- update -------------------------------------------------------------------
        pshift = 0;
        if (acc) {
            E = En;
            ++acc_total;
            if (nuis) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (q == pick) {
                        ++acc_r[q];
                        if (sgn != isg[q]) {
                            const int lo = since_sg[q] > burn ? since_sg[q] : burn;
                            if (s > lo) atomicAdd(g_param_counts + T.r[q].hist_sigma + isg[q], (unsigned long long)(s - lo));
                            since_sg[q] = s;
                            isg[q] = sgn;
                            cA[q] = An;
                            cD[q] = Dn;
                        }
                    }
                }
                if (HG && dgm != 0) {
                    const int lo = since_gm > burn ? since_gm : burn;
                    if (s > lo) atomicAdd(g_param_counts + T.r[QG].hist_gamma + igm, (unsigned long long)(s - lo));
                    since_gm = s;
                    igm = gmn;
                    pshift = dgm;
                    // slide the window and fetch the new edge
                    const double edge = __ldg(gsse + (size_t)state * G + wrap1(gmn + dgm, G));
                    if (dgm > 0) { W0 = W1; W1 = W2; W2 = edge; }
                    else { W2 = W1; W1 = W0; W0 = edge; }
                }
            } else {
                ++acc_state;
                if (nstate != state) {
                    const int lo = since_state > burn ? since_state : burn;
                    if (s > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s - lo));
                    since_state = s;
                    state = nstate;
                }
                Cst = Cn;
#pragma unroll
                for (int q = 0; q < R; ++q) { csse[q] = keep[q].x; cref[q] = keep[q].y; }
                if (HG) { W0 = nW0; W1 = num[QG]; W2 = nW2; }
            }
        }

        if (s >= burn) {
            if (A.state_trace) A.state_trace[(size_t)(A.s_begin + s - A.burn) * n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                A.traj_E[o] = E;
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
#pragma unroll
                for (int q = 0; q < R; ++q) A.traj_idx[((size_t)snap_idx * T.P + q) * n + c] = isg[q];
                if (HG) A.traj_idx[((size_t)snap_idx * T.P + R) * n + c] = igm;
                ++snap_idx;
                const long long ns = (long long)next_snap + A.traj_every;
                next_snap = ns < s_end ? (int)ns : -1;
            }
        }
        w = wn;
    }

    // ---- flush run-length counters, persist the chain ----------------------------------
    {
        const int lo = since_state > burn ? since_state : burn;
        if (s_end > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s_end - lo));
    }
    C.state[c] = state;
    C.E[c] = E;
    C.accepted[c] += acc_total;
    C.sep[(size_t)R * n + c] += acc_state;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int lo = since_sg[q] > burn ? since_sg[q] : burn;
        if (s_end > lo) atomicAdd(g_param_counts + T.r[q].hist_sigma + isg[q], (unsigned long long)(s_end - lo));
        C.isg[(size_t)q * n + c] = isg[q];
        C.sep[(size_t)q * n + c] += acc_r[q];
    }
    if (HG) {
        const int lo = since_gm > burn ? since_gm : burn;
        if (s_end > lo) atomicAdd(g_param_counts + T.r[QG].hist_gamma + igm, (unsigned long long)(s_end - lo));
        C.igm[(size_t)QG * n + c] = igm;
    }
}

template <int R, bool HG>
static cudaError_t launch_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid, int block,
                              cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_fast_kernel<R, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.sig_bytes);
    if (e != cudaSuccess) return e;
    mcmc_fast_kernel<R, HG><<<grid, block, T.sig_bytes, st>>>(T, C, A);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_fast(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid, int block,
                             cudaStream_t st) {
    if (!T.canonical) return cudaErrorNotSupported;
    const bool hg = T.r[T.R - 1].has_gamma != 0;
#define BCP_CASE(r)                                                                     \
    case r:                                                                             \
        return hg ? launch_one<r, true>(T, C, A, grid, block, st) : launch_one<r, false>(T, C, A, grid, block, st);
    switch (T.R) {
        BCP_CASE(1) BCP_CASE(2) BCP_CASE(3) BCP_CASE(4) BCP_CASE(5) BCP_CASE(6) BCP_CASE(7) BCP_CASE(8)
        default: return cudaErrorNotSupported;
    }
#undef BCP_CASE
}

}  // namespace bcp
