// K5 for ensembles that contain a Restraint_pf on the 6-D nuisance grid
// (/root/reference/biceps/Restraint.py:595-894).
//
// One WARP per chain.  Everything except the PF term is warp-uniform (every lane holds the same
// chain state and takes the same branches); the PF term
//     sse(i, cell)  = sum_j w_j (beta_c Nc_ij[xc,b] + beta_h Nh_ij[xh,b] + beta_0 - exp_j)^2      Restraint.py:742-749, 801
//     ref(i, cell)  = sum_j w_j max(-model_j, 0)          (exponential reference)                   Restraint.py:843-849
// is evaluated on the fly from the raw <N_c>/<N_h> counts: the 32 lanes stride over the residues
// (coalesced 856-byte rows) and the partial sums are combined by an xor butterfly.  That summation
// order (strided partials, then 16-8-4-2-1) is part of the specification: oracle/biceps_oracle.c:pf_eval
// does the same, so chains are bit-identical to the oracle.  The dense table the reference builds
// (11.6 MB per state for the default grid) is never materialised.
// All seven parameters of the PF restraint move together, like sigma and gamma of an NOE
// restraint (PosteriorSampler.py:303-305).
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

constexpr int PF_MAXR = BCP_MAX_RESTRAINTS;

__device__ __forceinline__ int pwrap1(int v, int n) {      // v in [-1, n]
    return v < 0 ? v + n : (v >= n ? v - n : v);
}

// warp-cooperative PF sums; every lane returns the same totals
__device__ __forceinline__ void pf_eval(const PfDev& pf, int state, const int* ci, int lane, double& sse, double& ref) {
    const int J = pf.J;
    const double bc = __ldg(pf.grid + ci[0]);
    const double bh = __ldg(pf.grid + pf.n[0] + ci[1]);
    const double b0 = __ldg(pf.grid + pf.n[0] + pf.n[1] + ci[2]);
    const double* __restrict__ Nc = pf.Nc + (((size_t)state * pf.n[3] + ci[3]) * pf.n[5] + ci[5]) * J;
    const double* __restrict__ Nh = pf.Nh + (((size_t)state * pf.n[4] + ci[4]) * pf.n[5] + ci[5]) * J;
    double ps = 0.0, pr = 0.0;
    for (int j = lane; j < J; j += 32) {
        const double model = __dadd_rn(__dadd_rn(__dmul_rn(bc, __ldg(Nc + j)), __dmul_rn(bh, __ldg(Nh + j))), b0);
        const double err = __dsub_rn(model, __ldg(pf.exp + j));
        const double w = __ldg(pf.weight + j);
        ps = __dadd_rn(ps, __dmul_rn(w, __dmul_rn(err, err)));
        const double neg = __dmul_rn(-1.0, model);
        pr = __dadd_rn(pr, __dmul_rn(w, neg > 0.0 ? neg : 0.0));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        ps = __dadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, off));
        pr = __dadd_rn(pr, __shfl_xor_sync(0xffffffffu, pr, off));
    }
    sse = ps;
    ref = pr;
}

__global__ void __launch_bounds__(128)
mcmc_pf_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
               const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);

    const long long n = C.n_chains;
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;                                   // whole warps only
    const bool lead = lane == 0;
    const int R = T.R;
    const int S = T.S;
    const PfDev& pf = T.pf;
    const int QP = pf.q;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * S;
    const double logZ = T.logZ[lam];
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
    const uint32_t thresh = T.nuis_thresh;

    // restraint term in the reference's order: ((nlsig + sse/two_sig2) [+ prior]) + cnorm [- ref]
    auto term = [&](int q, int st, int sg, int gm, const int* ci) -> double {
        const RestraintDev& rd = T.r[q];
        const double a = s_sig[rd.sig_off + sg];
        const double d = s_sig[rd.sig_off + rd.n_sigma + sg];
        if (q == QP) {
            double sse, ref;
            pf_eval(pf, st, ci, lane, sse, ref);
            double t = __dadd_rn(a, __ddiv_rn(sse, d));
            if (pf.prior) {
                size_t cell = 0;
#pragma unroll
                for (int k = 0; k < BCP_PF_NGRID; ++k) cell = cell * (size_t)pf.n[k] + (size_t)ci[k];
                t = __dadd_rn(t, __ldg(pf.prior + cell));
            }
            t = __dadd_rn(t, rd.cnorm);
            if (pf.ref_kind == 1) t = __dsub_rn(t, ref);
            return t;
        }
        const double sse = __ldg(rd.sse + (size_t)st * rd.n_gamma + gm);
        double t = __dadd_rn(a, __ddiv_rn(sse, d));
        t = __dadd_rn(t, rd.cnorm);
        if (rd.refsum) t = __dsub_rn(t, __ldg(rd.refsum + st));
        return t;
    };

    int state = C.state[c];
    double E = C.E[c];
    int isg[PF_MAXR], igm[PF_MAXR], ipf[BCP_PF_NGRID];
    double t[PF_MAXR];
    long long since_sg[PF_MAXR], since_gm[PF_MAXR], since_pf[BCP_PF_NGRID];
    unsigned int acc_r[PF_MAXR];
    unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
    for (int k = 0; k < BCP_PF_NGRID; ++k) { ipf[k] = QP >= 0 ? C.ipf[(size_t)k * n + c] : 0; since_pf[k] = A.s_begin; }
#pragma unroll
    for (int q = 0; q < PF_MAXR; ++q) {
        isg[q] = 0; igm[q] = 0; t[q] = 0.0; since_sg[q] = A.s_begin; since_gm[q] = A.s_begin; acc_r[q] = 0;
        if (q < R) {
            isg[q] = C.isg[(size_t)q * n + c];
            igm[q] = C.igm[(size_t)q * n + c];
            t[q] = term(q, state, isg[q], igm[q], ipf);
        }
    }
    double Cst = __dadd_rn(__ldg(energy + state), logZ);
    long long since_state = A.s_begin;
    long long next_snap = -1, snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        next_snap = A.burn + m * A.traj_every;
    }
    auto flush = [&](unsigned long long* bin, long long since, long long s) {
        const long long lo = since > A.burn ? since : A.burn;
        if (lead && s > lo) atomicAdd(bin, (unsigned long long)(s - lo));
    };

    for (long long s = A.s_begin; s < A.s_end; ++s) {
        const Philox4 w = step_block(C.seed, chain_id, A.step0 + (unsigned long long)s);
        const bool nuis = w.w0 < thresh;
        const uint64_t pr = (uint64_t)w.w1 * (uint32_t)(nuis ? R : S);
        const int hi = (int)(pr >> 32);
        uint32_t f = (uint32_t)pr;
        const int pick = nuis ? hi : -1;
        const int nstate = nuis ? state : hi;
        auto delta = [&]() {
            const uint64_t z = (uint64_t)f * 3u;
            f = (uint32_t)z;
            return (int)(z >> 32) - 1;
        };
        int nsg = 0, ngm = 0, npf[BCP_PF_NGRID];
#pragma unroll
        for (int k = 0; k < BCP_PF_NGRID; ++k) npf[k] = ipf[k];
        double tn[PF_MAXR];
        double En = nuis ? Cst : __dadd_rn(__ldg(energy + nstate), logZ);
        const double Cn = En;
#pragma unroll
        for (int q = 0; q < PF_MAXR; ++q) {
            if (q < R) {
                tn[q] = t[q];
                if (!nuis) {
                    tn[q] = term(q, nstate, isg[q], igm[q], ipf);
                } else if (q == pick) {                                   // warp-uniform branch
                    nsg = pwrap1(isg[q] + delta(), T.r[q].n_sigma);
                    ngm = igm[q];
                    if (T.r[q].has_gamma) ngm = pwrap1(igm[q] + delta(), T.r[q].n_gamma);
                    if (q == QP) {
#pragma unroll
                        for (int k = 0; k < BCP_PF_NGRID; ++k) npf[k] = pwrap1(ipf[k] + delta(), pf.n[k]);
                    }
                    tn[q] = term(q, state, nsg, ngm, npf);
                }
                En = __dadd_rn(En, tn[q]);
            }
        }
        bool acc;
        if (En < E) acc = true;
        else acc = metropolis_accept(__dsub_rn(E, En), u2_from_words(w.w2, w.w3));

        if (acc) {
            E = En;
            ++acc_total;
            if (!nuis) {
                ++acc_state;
                Cst = Cn;
                if (nstate != state) {
                    flush(g_state_counts + state, since_state, s);
                    since_state = s;
                    state = nstate;
                }
            }
#pragma unroll
            for (int q = 0; q < PF_MAXR; ++q) {
                if (q < R) {
                    t[q] = tn[q];
                    if (nuis && q == pick) {
                        ++acc_r[q];
                        if (nsg != isg[q]) {
                            flush(g_param_counts + T.r[q].hist_sigma + isg[q], since_sg[q], s);
                            since_sg[q] = s;
                            isg[q] = nsg;
                        }
                        if (ngm != igm[q]) {
                            flush(g_param_counts + T.r[q].hist_gamma + igm[q], since_gm[q], s);
                            since_gm[q] = s;
                            igm[q] = ngm;
                        }
                        if (q == QP) {
#pragma unroll
                            for (int k = 0; k < BCP_PF_NGRID; ++k)
                                if (npf[k] != ipf[k]) {
                                    flush(g_param_counts + pf.hist[k] + ipf[k], since_pf[k], s);
                                    since_pf[k] = s;
                                    ipf[k] = npf[k];
                                }
                        }
                    }
                }
            }
        }

        if (s >= A.burn && lead) {
            if (A.state_trace && c < A.trace_n) A.state_trace[(size_t)(s - A.burn) * A.trace_n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                A.traj_E[o] = E;
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
                int p = 0;
#pragma unroll
                for (int q = 0; q < PF_MAXR; ++q) {
                    if (q < R) {
                        A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = isg[q];
                        if (T.r[q].has_gamma) A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = igm[q];
                        if (q == QP) {
#pragma unroll
                            for (int k = 0; k < BCP_PF_NGRID; ++k) A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = ipf[k];
                        }
                    }
                }
            }
        }
        if (s >= A.burn && s == next_snap) { ++snap_idx; next_snap += A.traj_every; }
    }

    flush(g_state_counts + state, since_state, A.s_end);
#pragma unroll
    for (int q = 0; q < PF_MAXR; ++q) {
        if (q < R) {
            flush(g_param_counts + T.r[q].hist_sigma + isg[q], since_sg[q], A.s_end);
            if (T.r[q].has_gamma) flush(g_param_counts + T.r[q].hist_gamma + igm[q], since_gm[q], A.s_end);
            if (lead) {
                C.isg[(size_t)q * n + c] = isg[q];
                C.igm[(size_t)q * n + c] = igm[q];
                C.sep[(size_t)q * n + c] += acc_r[q];
            }
        }
    }
    if (QP >= 0) {
#pragma unroll
        for (int k = 0; k < BCP_PF_NGRID; ++k) {
            flush(g_param_counts + pf.hist[k] + ipf[k], since_pf[k], A.s_end);
            if (lead) C.ipf[(size_t)k * n + c] = ipf[k];
        }
    }
    if (lead) {
        C.state[c] = state;
        C.E[c] = E;
        C.accepted[c] += acc_total;
        C.sep[(size_t)R * n + c] += acc_state;
    }
}

// -log P of explicit configurations for ensembles with a PFGRID restraint, one warp per configuration
// (PosteriorSampler.neglogP order).  all_lambdas: out[l][i] for every lambda l (u_kn), else out[i] for lam[i].
__global__ void __launch_bounds__(128)
energy_pf_kernel(const SamplerTables T, long long N, const int* __restrict__ lam, const int* __restrict__ state,
                 const int* __restrict__ indices, double* __restrict__ out, int all_lambdas) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const int st = state[i];
    double t[PF_MAXR];
    int p = 0;
#pragma unroll
    for (int q = 0; q < PF_MAXR; ++q) {
        t[q] = 0.0;
        if (q < T.R) {
            const RestraintDev& rd = T.r[q];
            const int sg = indices[(size_t)i * T.P + p++];
            const double* sig = T.sig_tab + rd.sig_off;
            double v;
            if (q == T.pf.q) {
                int ci[BCP_PF_NGRID];
#pragma unroll
                for (int k = 0; k < BCP_PF_NGRID; ++k) ci[k] = indices[(size_t)i * T.P + p++];
                double sse, ref;
                pf_eval(T.pf, st, ci, lane, sse, ref);
                v = __dadd_rn(sig[sg], __ddiv_rn(sse, sig[rd.n_sigma + sg]));
                if (T.pf.prior) {
                    size_t cell = 0;
#pragma unroll
                    for (int k = 0; k < BCP_PF_NGRID; ++k) cell = cell * (size_t)T.pf.n[k] + (size_t)ci[k];
                    v = __dadd_rn(v, T.pf.prior[cell]);
                }
                v = __dadd_rn(v, rd.cnorm);
                if (T.pf.ref_kind == 1) v = __dsub_rn(v, ref);
            } else {
                const int gm = rd.has_gamma ? indices[(size_t)i * T.P + p++] : 0;
                v = __dadd_rn(sig[sg], __ddiv_rn(rd.sse[(size_t)st * rd.n_gamma + gm], sig[rd.n_sigma + sg]));
                v = __dadd_rn(v, rd.cnorm);
                if (rd.refsum) v = __dsub_rn(v, rd.refsum[st]);
            }
            t[q] = v;
        }
    }
    if (lane != 0) return;
    const int l0 = all_lambdas ? 0 : (lam ? lam[i] : 0);
    const int l1 = all_lambdas ? T.K : l0 + 1;
    for (int l = l0; l < l1; ++l) {
        double E = __dadd_rn(T.energy[(size_t)l * T.S + st], T.logZ[l]);
#pragma unroll
        for (int q = 0; q < PF_MAXR; ++q)
            if (q < T.R) E = __dadd_rn(E, t[q]);
        out[all_lambdas ? (size_t)l * N + i : (size_t)i] = E;
    }
}

cudaError_t launch_energy_pf(const SamplerTables& T, long long N, const int* d_lam, const int* d_state,
                             const int* d_indices, double* d_out, int all_lambdas, cudaStream_t st) {
    const long long threads = N * 32;
    energy_pf_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(T, N, d_lam, d_state, d_indices, d_out, all_lambdas);
    return cudaGetLastError();
}

cudaError_t launch_mcmc_pf(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_pf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T.sig_bytes);
    if (e != cudaSuccess) return e;
    const int block = 128;                                   // 4 chains per CTA
    const long long threads = C.n_chains * 32;
    const int grid = (int)((threads + block - 1) / block);
    mcmc_pf_kernel<<<grid, block, T.sig_bytes, st>>>(T, C, A);
    return cudaGetLastError();
}

}  // namespace bcp
