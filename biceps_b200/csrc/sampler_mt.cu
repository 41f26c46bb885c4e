// K5, reference-stream mode: ONE chain driven by numpy's legacy MT19937 stream, draw for draw like the
// reference's own sampler, so that np.random.seed(s) gives the reference's trajectory bit for bit on the GPU.
//
// The reference consumes a data-dependent number of variates per step from numpy's global generator
// (/root/reference/biceps/PosteriorSampler.py:298-307, 322-326):
//     dice = random()                                    move type: nuisance iff dice < 1 - 1/(R+1)
//     nuisance move: randint(R), then randint(3) - 1 for every parameter of that restraint
//     state move   : randint(low=0, high=S, size=1)
//     accept       : E_new < E, else random() < exp(E - E_new)   (the second uniform only then)
// The host hands over the raw 32-bit outputs of MT19937 (RandomState._bit_generator.random_raw) and this
// kernel replays numpy's legacy algorithms on them:
//     random()    = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two outputs                 (mt19937_next_double)
//     randint(n)  = masked rejection: draw & mask until <= n - 1; n == 1 draws nothing  (buffered_bounded_masked_uint32)
// and reports how many outputs it consumed so that the caller can advance the global generator by the same
// amount.  Single thread, any restraint layout (PF-grid restraints included: their term is summed over the residues in
// the reference's own order here); this is a compatibility / validation path
// (tests/test_mt_stream_gpu.py replays the reference's stored trajectories), not a performance path.
#include "common.cuh"
#include "sampler_common.cuh"

#include <math.h>

namespace bcp {

struct MtCursor {
    const uint32_t* raw;
    long long len, pos;
    bool dry;
    __device__ uint32_t next() {
        if (pos >= len) { dry = true; return 0u; }
        return raw[pos++];
    }
    __device__ double uniform() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    __device__ int randint(int n) {
        const uint32_t rng = (uint32_t)(n - 1);
        if (rng == 0u) return 0;
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next() & mask; } while (v > rng && !dry);
        return (int)v;
    }
};

// restraint term in the reference's order: ((nlsig + sse / two_sig2) [+ prior]) + cnorm [- ref].  The PF-grid term is
// summed over the residues left to right, which is the order in which the reference accumulates its dense 6-D
// tables (`sse += weight * err**2.0` per residue, Restraint.py:742-749; `sum += weight * max(-model, 0)`, :843-849;
// model = beta_c * Nc + beta_h * Nh + beta_0 elementwise, :765-801) -- so energies equal the reference's bit for bit.
__device__ double mt_term(const SamplerTables& T, const double* s_sig, int q, int state, int sg, int gm, const int* ci) {
    const RestraintDev& rd = T.r[q];
    const double a = s_sig[rd.sig_off + sg], d = s_sig[rd.sig_off + rd.n_sigma + sg];
    if (q == T.pf.q) {
        const PfDev& pf = T.pf;
        const int J = pf.J;
        const double bc = pf.grid[ci[0]], bh = pf.grid[pf.n[0] + ci[1]], b0 = pf.grid[pf.n[0] + pf.n[1] + ci[2]];
        const double* Nc = pf.Nc + (((size_t)state * pf.n[3] + ci[3]) * pf.n[5] + ci[5]) * J;
        const double* Nh = pf.Nh + (((size_t)state * pf.n[4] + ci[4]) * pf.n[5] + ci[5]) * J;
        double sse = 0.0, ref = 0.0;
        for (int j = 0; j < J; ++j) {
            const double model = __dadd_rn(__dadd_rn(__dmul_rn(bc, Nc[j]), __dmul_rn(bh, Nh[j])), b0);
            const double err = __dsub_rn(model, pf.exp[j]);
            sse = __dadd_rn(sse, __dmul_rn(pf.weight[j], __dmul_rn(err, err)));
            const double neg = __dmul_rn(-1.0, model);
            ref = __dadd_rn(ref, __dmul_rn(pf.weight[j], neg > 0.0 ? neg : 0.0));
        }
        double t = __dadd_rn(a, __ddiv_rn(sse, d));
        if (pf.prior) {
            size_t cell = 0;
            for (int k = 0; k < BCP_PF_NGRID; ++k) cell = cell * (size_t)pf.n[k] + (size_t)ci[k];
            t = __dadd_rn(t, pf.prior[cell]);
        }
        t = __dadd_rn(t, rd.cnorm);
        if (pf.ref_kind == 1) t = __dsub_rn(t, ref);
        return t;
    }
    const double sse = __ldg(rd.sse + (size_t)state * rd.n_gamma + gm);
    double t = __dadd_rn(a, __ddiv_rn(sse, d));
    t = __dadd_rn(t, rd.cnorm);
    if (rd.refsum) t = __dsub_rn(t, __ldg(rd.refsum + state));
    return t;
}

// R is a template parameter so that the per-restraint state (indices, terms, counters) lives in registers: every
// access below uses a compile-time index (`q == pick` selects instead of `isg[pick]`)
template <int R>
__global__ void __launch_bounds__(32)
mcmc_mt_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
               const __grid_constant__ LaunchArgs A, MtStream M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);
    if (threadIdx.x != 0) return;

    const long long n = C.n_chains, c = 0;
    const int lam = C.lam[c];
    const int grp = C.group[c];
    const double* energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
    MtCursor rng{M.raw, M.len, *M.pos, false};
    const double p_nuis = 1. - 1. / ((double)R + 1.);            // PosteriorSampler.py:300

    int state = C.state[c];
    double E = C.E[c];
    int isg[R], igm[R], ipf[BCP_PF_NGRID];
    double t[R];
    unsigned int acc_r[R];
    unsigned int acc_state = 0, acc_total = 0;
    const int QP = T.pf.q;
#pragma unroll
    for (int k = 0; k < BCP_PF_NGRID; ++k) ipf[k] = QP >= 0 ? C.ipf[(size_t)k * n + c] : 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        isg[q] = C.isg[(size_t)q * n + c];
        igm[q] = C.igm[(size_t)q * n + c];
        t[q] = mt_term(T, s_sig, q, state, isg[q], igm[q], ipf);
        acc_r[q] = 0;
    }
    double Cst = __dadd_rn(__ldg(energy + state), logZ);
    long long next_snap = -1, snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        next_snap = A.burn + m * A.traj_every;
    }

    for (long long s = A.s_begin; s < A.s_end; ++s) {
        const bool nuis = rng.uniform() < p_nuis;
        int pick = -1, nstate = state, nsg = 0, ngm = 0, npf[BCP_PF_NGRID];
        double tn[R];
#pragma unroll
        for (int q = 0; q < R; ++q) tn[q] = t[q];
#pragma unroll
        for (int k = 0; k < BCP_PF_NGRID; ++k) npf[k] = ipf[k];
        if (nuis) {
            pick = rng.randint(R);
            int cur_sg = 0, cur_gm = 0;
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (q == pick) { cur_sg = isg[q]; cur_gm = igm[q]; }
            const RestraintDev& rd = T.r[pick];
            int v = cur_sg + rng.randint(3) - 1;
            nsg = v < 0 ? v + rd.n_sigma : (v >= rd.n_sigma ? v - rd.n_sigma : v);
            ngm = cur_gm;
            if (rd.has_gamma) {
                v = cur_gm + rng.randint(3) - 1;
                ngm = v < 0 ? v + rd.n_gamma : (v >= rd.n_gamma ? v - rd.n_gamma : v);
            }
            if (pick == QP) {              // the six grid parameters follow sigma, in the reference's parameter order
#pragma unroll
                for (int k = 0; k < BCP_PF_NGRID; ++k) {
                    v = ipf[k] + rng.randint(3) - 1;
                    npf[k] = v < 0 ? v + T.pf.n[k] : (v >= T.pf.n[k] ? v - T.pf.n[k] : v);
                }
            }
            const double tp = mt_term(T, s_sig, pick, state, nsg, ngm, npf);
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (q == pick) tn[q] = tp;
        } else {
            nstate = rng.randint(T.S);
#pragma unroll
            for (int q = 0; q < R; ++q) tn[q] = mt_term(T, s_sig, q, nstate, isg[q], igm[q], ipf);
        }
        const double Cn = nuis ? Cst : __dadd_rn(__ldg(energy + nstate), logZ);
        double En = Cn;
#pragma unroll
        for (int q = 0; q < R; ++q) En = __dadd_rn(En, tn[q]);

        bool acc = En < E;
        if (!acc) acc = rng.uniform() < exp(__dsub_rn(E, En));   // PosteriorSampler.py:325
        if (rng.dry) break;

        if (acc) {
            E = En;
            if (!nuis) {
                state = nstate;
                Cst = Cn;
                ++acc_state;
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (q == pick) { ++acc_r[q]; isg[q] = nsg; igm[q] = ngm; }
#pragma unroll
                for (int k = 0; k < BCP_PF_NGRID; ++k) ipf[k] = npf[k];
            }
#pragma unroll
            for (int q = 0; q < R; ++q) t[q] = tn[q];
            ++acc_total;
        }
        if (s >= A.burn) {
            // one chain: plain per-step counting, no run-length compression needed
            atomicAdd(g_state_counts + state, 1ull);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                atomicAdd(g_param_counts + T.r[q].hist_sigma + isg[q], 1ull);
                if (T.r[q].has_gamma) atomicAdd(g_param_counts + T.r[q].hist_gamma + igm[q], 1ull);
            }
            if (QP >= 0)
                for (int k = 0; k < BCP_PF_NGRID; ++k) atomicAdd(g_param_counts + T.pf.hist[k] + ipf[k], 1ull);
            if (A.state_trace) A.state_trace[(size_t)(s - A.burn) * A.trace_n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                A.traj_E[o] = E;
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
                int p = 0;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = isg[q];
                    if (T.r[q].has_gamma) A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = igm[q];
                    if (q == QP)
                        for (int k = 0; k < BCP_PF_NGRID; ++k) A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = ipf[k];
                }
                ++snap_idx;
                next_snap += A.traj_every;
            }
        }
    }
    C.state[c] = state;
    C.E[c] = E;
    C.accepted[c] += acc_total;
    C.sep[(size_t)R * n + c] += acc_state;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        C.isg[(size_t)q * n + c] = isg[q];
        C.igm[(size_t)q * n + c] = igm[q];
        C.sep[(size_t)q * n + c] += acc_r[q];
    }
    if (QP >= 0)
        for (int k = 0; k < BCP_PF_NGRID; ++k) C.ipf[(size_t)k * n + c] = ipf[k];
    *M.pos = rng.pos;
    if (rng.dry) *M.dry = 1;
}

template <int R>
static cudaError_t launch_mt_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, const MtStream& M,
                                 cudaStream_t st) {
    const size_t smem = (size_t)T.sig_bytes + 16;
    cudaError_t e = cudaFuncSetAttribute(mcmc_mt_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mcmc_mt_kernel<R><<<1, 32, smem, st>>>(T, C, A, M);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_mt(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, const MtStream& M,
                           cudaStream_t st) {
    if (C.n_chains != 1) return cudaErrorNotSupported;
    switch (T.R) {
        case 1: return launch_mt_one<1>(T, C, A, M, st);
        case 2: return launch_mt_one<2>(T, C, A, M, st);
        case 3: return launch_mt_one<3>(T, C, A, M, st);
        case 4: return launch_mt_one<4>(T, C, A, M, st);
        case 5: return launch_mt_one<5>(T, C, A, M, st);
        case 6: return launch_mt_one<6>(T, C, A, M, st);
        case 7: return launch_mt_one<7>(T, C, A, M, st);
        case 8: return launch_mt_one<8>(T, C, A, M, st);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace bcp
