// K5-PF, tuned: one warp per chain for ensembles whose restraints are SCALAR ones plus one Restraint_pf on
// the 6-D nuisance grid (the apomyoglobin shape and BASELINE configs[3]; anything else runs sampler_pf.cu).
//
// The PF term is evaluated on the fly from two 8 J-byte rows <N_c>[state][x_c][b][:], <N_h>[state][x_h][b][:]
// (sampler_pf.cu).  At C4 size (10^5 states, 1.3 GB of counts) those rows come from HBM at random, 16 J
// bytes per evaluation: this kernel is HBM bound (6554 GB/s / 1712 B = 3.8e9 evaluations/s), and what it
// needs is enough warps in flight, not fewer instructions:
//   * R and the residues per lane (NJ = ceil(J / 32) <= 4) are template parameters, the chain state is a
//     handful of scalars: 2-3x fewer registers than the generic kernel's 235 -> 4+ warps per sub-partition;
//   * the rows of step s+1's proposal are requested while step s is still being evaluated, assuming step s
//     is rejected (PF acceptance is a few per cent); an accepted step simply re-requests them.  They travel
//     by cp.async into a per-warp double buffer in shared memory, so data in flight holds no registers;
//   * lane j generates the Philox block of step s0 + j for a batch of 32 steps (1/32 block per step and
//     lane instead of a whole one).
// The summation order of the PF sums (strided partials per lane, then the 16-8-4-2-1 xor butterfly) and
// every fp64 operation are those of sampler_pf.cu / oracle/biceps_oracle.c:pf_eval, so chains stay
// bit-identical to the oracle.
#include "exp_det.cuh"
#include "philox.cuh"
#include "sampler_common.cuh"

namespace bcp {

__device__ __forceinline__ int p2wrap(int v, int n) {      // v in [-1, n]
    return v < 0 ? v + n : (v >= n ? v - n : v);
}

// resident CTAs per SM asked of ptxas: 4 (128 registers) while the chain's scalar restraints fit, 3 (168 registers) from
// three scalar restraints on -- the reference's apomyoglobin shape (cs_H, cs_Ca, cs_N + PF) spilled 40-176 bytes at 128
#ifndef PF2_MIN_BLOCKS
#define PF2_MIN_BLOCKS(R) ((R) <= 2 ? 4 : 3)
#endif

struct Pf2Prop {            // proposal of one step, decoded from its Philox block (warp-uniform)
    bool nuis;
    int pick;               // restraint that moves (nuisance move)
    int nstate;             // proposed state (state move)
    int nsg;                // proposed sigma index of `pick`
    int npf[BCP_PF_NGRID];  // proposed PF grid indices (pick == PF restraint)
};

template <int R, int NJ>
__global__ void __launch_bounds__(128, PF2_MIN_BLOCKS(R))
mcmc_pf2_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);
    // experimental values and weights of the residues, padded with zero weight up to 32 NJ
    double* s_ex = s_sig + T.sig_bytes / 8;
    double* s_wt = s_ex + 32 * NJ;
    double* s_rows = s_wt + 32 * NJ;                        // [warps][2 buffers][nc | nh][32 NJ], zero past J
    for (int k = threadIdx.x; k < 32 * NJ; k += blockDim.x) {
        s_ex[k] = k < T.pf.J ? T.pf.exp[k] : 0.0;
        s_wt[k] = k < T.pf.J ? T.pf.weight[k] : 0.0;
    }
    for (int k = threadIdx.x; k < (int)(blockDim.x >> 5) * 4 * 32 * NJ; k += blockDim.x) s_rows[k] = 0.0;
    __syncthreads();

    const long long n = C.n_chains;
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;                                   // whole warps only
    const bool lead = lane == 0;
    const int S = T.S;
    const PfDev& pf = T.pf;
    const int QP = pf.q;
    const int J = pf.J;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * S;
    const double logZ = T.logZ[lam];
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
    const uint32_t thresh = T.nuis_thresh;
    const int s_end = (int)(A.s_end - A.s_begin);
    const long long burn_l = A.burn - A.s_begin;
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;

    const int n0 = pf.n[0], n1 = pf.n[1], n2 = pf.n[2], n3 = pf.n[3], n4 = pf.n[4], n5 = pf.n[5];
    const int nk[BCP_PF_NGRID] = {n0, n1, n2, n3, n4, n5};
    // ---- the two rows of counts of (state, grid indices), and the PF sums over them ----------------
    double* my_rows = s_rows + (size_t)(threadIdx.x >> 5) * 4 * 32 * NJ;
    const uint32_t my_rows_u32 = smem_u32(my_rows);
    // cp.async of the two rows into buffer `buf`; each lane copies (and later reads) its own residues
    auto request_rows = [&](int st, const int* ci, int buf) {
        const double* __restrict__ Nc = pf.Nc + (size_t)((unsigned)((st * n3 + ci[3]) * n5 + ci[5])) * (unsigned)J;
        const double* __restrict__ Nh = pf.Nh + (size_t)((unsigned)((st * n4 + ci[4]) * n5 + ci[5])) * (unsigned)J;
        const uint32_t dst = my_rows_u32 + (uint32_t)buf * (2u * 32u * NJ * 8u) + 8u * (uint32_t)lane;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
            const int j = lane + 32 * i;
            if (j < J) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 256u * i), "l"(Nc + j) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 256u * (NJ + i)), "l"(Nh + j) : "memory");
            }
        }
    };
    auto pf_sums = [&](const int* ci, int buf, double& sse, double& ref) {
        const double* nc = my_rows + (size_t)buf * (2 * 32 * NJ) + lane;
        const double* nh = nc + 32 * NJ;
        const double bc = __ldg(pf.grid + ci[0]);
        const double bh = __ldg(pf.grid + n0 + ci[1]);
        const double b0 = __ldg(pf.grid + n0 + n1 + ci[2]);
        double ps = 0.0, pr = 0.0;
        // residues past J carry weight 0 and counts 0: they add +0.0, which changes no bit of ps / pr
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
            const double model = __dadd_rn(__dadd_rn(__dmul_rn(bc, nc[32 * i]), __dmul_rn(bh, nh[32 * i])), b0);
            const double err = __dsub_rn(model, s_ex[lane + 32 * i]);
            const double wi = s_wt[lane + 32 * i];
            ps = __dadd_rn(ps, __dmul_rn(wi, __dmul_rn(err, err)));
            // reference: pr += w * max(-model, 0).  For model >= 0 (or -0.0, NaN) that adds w * 0.0 = +0.0, which
            // leaves every bit of pr unchanged (pr is never -0.0), so the add is skipped
            if (model < 0.0) pr = __dadd_rn(pr, __dmul_rn(wi, __dmul_rn(-1.0, model)));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            ps = __dadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, off));
            pr = __dadd_rn(pr, __shfl_xor_sync(0xffffffffu, pr, off));
        }
        sse = ps;
        ref = pr;
    };
    // restraint terms in the reference's order: ((nlsig + sse/two_sig2) [+ prior]) + cnorm [- ref]
    auto pf_term = [&](int sg, const int* ci, double sse, double ref) {
        const RestraintDev& rd = T.r[QP];
        double t = __dadd_rn(s_sig[rd.sig_off + sg], __ddiv_rn(sse, s_sig[rd.sig_off + rd.n_sigma + sg]));
        if (pf.prior) {
            size_t cell = 0;
#pragma unroll
            for (int k = 0; k < BCP_PF_NGRID; ++k) cell = cell * (size_t)nk[k] + (size_t)ci[k];
            t = __dadd_rn(t, __ldg(pf.prior + cell));
        }
        t = __dadd_rn(t, rd.cnorm);
        if (pf.ref_kind == 1) t = __dsub_rn(t, ref);
        return t;
    };
    auto scalar_term = [&](int q, int sg, double sse, double refv) {
        const RestraintDev& rd = T.r[q];
        double t = __dadd_rn(s_sig[rd.sig_off + sg], __ddiv_rn(sse, s_sig[rd.sig_off + rd.n_sigma + sg]));
        t = __dadd_rn(t, rd.cnorm);
        if (rd.refsum) t = __dsub_rn(t, refv);
        return t;
    };

    // ---- chain state (warp-uniform, replicated in every lane) ---------------------------------
    int state = C.state[c];
    double E = C.E[c];
    int isg[R], ipf[BCP_PF_NGRID];
    double t[R], csse[R], cref[R];
    int since_sg[R], since_pf[BCP_PF_NGRID], since_state = burn;
    unsigned int acc_r[R];
    unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
    for (int k = 0; k < BCP_PF_NGRID; ++k) { ipf[k] = C.ipf[(size_t)k * n + c]; since_pf[k] = burn; }
    {
        request_rows(state, ipf, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int q = 0; q < R; ++q) {
            isg[q] = C.isg[(size_t)q * n + c];
            since_sg[q] = burn;
            acc_r[q] = 0;
            csse[q] = 0.0; cref[q] = 0.0;
            if (q == QP) {
                double sse, ref;
                pf_sums(ipf, 0, sse, ref);
                t[q] = pf_term(isg[q], ipf, sse, ref);
            } else {
                csse[q] = __ldg(T.r[q].sse + state);
                cref[q] = T.r[q].refsum ? __ldg(T.r[q].refsum + state) : 0.0;
                t[q] = scalar_term(q, isg[q], csse[q], cref[q]);
            }
        }
    }
    double Cst = __dadd_rn(__ldg(energy + state), logZ);
    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }
    auto flush = [&](unsigned long long* bin, int since, int s) {      // run length of the value that is left
        if (lead && s > burn && s > since) atomicAdd(bin, (unsigned long long)(s - since));
    };

    // ---- proposals: Philox blocks of 32 steps, one per lane ------------------------------------------
    const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
    const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);
    auto block_of = [&](int s) {
        const unsigned long long ctr = ctr0 + (unsigned long long)s;
        return philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), ch_lo, ch_hi, seed_lo, seed_hi);
    };
    auto bcast = [&](const Philox4& b, int j) {
        Philox4 r;
        r.w0 = __shfl_sync(0xffffffffu, b.w0, j); r.w1 = __shfl_sync(0xffffffffu, b.w1, j);
        r.w2 = __shfl_sync(0xffffffffu, b.w2, j); r.w3 = __shfl_sync(0xffffffffu, b.w3, j);
        return r;
    };
    auto decode = [&](const Philox4& w) {                               // on the CURRENT chain state
        Pf2Prop p;
        p.nuis = w.w0 < thresh;
        const uint64_t pr = (uint64_t)w.w1 * (uint32_t)(p.nuis ? R : S);
        const int hi = (int)(pr >> 32);
        uint32_t f = (uint32_t)pr;
        p.pick = p.nuis ? hi : -1;
        p.nstate = p.nuis ? state : hi;
        p.nsg = 0;
#pragma unroll
        for (int k = 0; k < BCP_PF_NGRID; ++k) p.npf[k] = ipf[k];
        if (p.nuis) {
            uint64_t z = (uint64_t)f * 3u;
            f = (uint32_t)z;
            int sgq = 0, nsq = 1;
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (q == hi) { sgq = isg[q]; nsq = T.r[q].n_sigma; }
            p.nsg = p2wrap(sgq + (int)(z >> 32) - 1, nsq);
            if (hi == QP) {
#pragma unroll
                for (int k = 0; k < BCP_PF_NGRID; ++k) {
                    z = (uint64_t)f * 3u;
                    f = (uint32_t)z;
                    p.npf[k] = p2wrap(ipf[k] + (int)(z >> 32) - 1, nk[k]);
                }
            }
        }
        return p;
    };
    // what a proposal reads from global memory (requested one step ahead)
    struct Pre { double C, sse[R], ref[R]; };
    auto request = [&](const Pf2Prop& p, Pre& x, int buf) {
        if (!p.nuis || p.pick == QP) request_rows(p.nstate, p.npf, buf);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (!p.nuis) {
            x.C = __ldg(energy + p.nstate);
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (q != QP) {
                    x.sse[q] = __ldg(T.r[q].sse + p.nstate);
                    x.ref[q] = T.r[q].refsum ? __ldg(T.r[q].refsum + p.nstate) : 0.0;
                }
        }
    };

    Philox4 gcur = block_of(lane), gnxt = block_of(32 + lane);

    // one step: (w, prop, pre) describe step s and are consumed; (wn, pn, pren) are produced for step s+1.
    // The caller alternates two sets of buffers, so nothing is copied.
    auto step = [&](const int s, const Philox4& w, const Pf2Prop& prop, const Pre& pre, Philox4& wn, Pf2Prop& pn, Pre& pren) {
        const int j = s & 31;
        // the next step's block; its proposal is decoded on the current state (= assuming step s is rejected)
        // and its rows are requested before step s is evaluated
        wn = (j == 31) ? bcast(gnxt, 0) : bcast(gcur, j + 1);
        if (j == 31) { gcur = gnxt; gnxt = block_of(s + 1 + 32 + lane); }
        pn = decode(wn);
        request(pn, pren, (s + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // step s's rows have landed

        double tn[R];
        const double Cn = prop.nuis ? Cst : __dadd_rn(pre.C, logZ);
        double En = Cn;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            tn[q] = t[q];
            if (!prop.nuis || q == prop.pick) {                            // warp-uniform
                const int sg = prop.nuis ? prop.nsg : isg[q];
                if (q == QP) {
                    double sse, ref;
                    pf_sums(prop.npf, s & 1, sse, ref);
                    tn[q] = pf_term(sg, prop.npf, sse, ref);
                } else {
                    tn[q] = scalar_term(q, sg, prop.nuis ? csse[q] : pre.sse[q], prop.nuis ? cref[q] : pre.ref[q]);
                }
            }
            En = __dadd_rn(En, tn[q]);
        }
        bool acc;
        if (En < E) acc = true;
        else acc = metropolis_accept(__dsub_rn(E, En), u2_from_words(w.w2, w.w3));

        if (acc) {
            E = En;
            ++acc_total;
            if (!prop.nuis) {
                ++acc_state;
                Cst = Cn;
                if (prop.nstate != state) {
                    flush(g_state_counts + state, since_state, s);
                    if (s > burn) since_state = s;
                    state = prop.nstate;
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                t[q] = tn[q];
                if (!prop.nuis && q != QP) { csse[q] = pre.sse[q]; cref[q] = pre.ref[q]; }
                if (prop.nuis && q == prop.pick) {
                    ++acc_r[q];
                    if (prop.nsg != isg[q]) {
                        flush(g_param_counts + T.r[q].hist_sigma + isg[q], since_sg[q], s);
                        if (s > burn) since_sg[q] = s;
                        isg[q] = prop.nsg;
                    }
                    if (q == QP) {
#pragma unroll
                        for (int k = 0; k < BCP_PF_NGRID; ++k)
                            if (prop.npf[k] != ipf[k]) {
                                flush(g_param_counts + pf.hist[k] + ipf[k], since_pf[k], s);
                                if (s > burn) since_pf[k] = s;
                                ipf[k] = prop.npf[k];
                            }
                    }
                }
            }
            // the request for step s+1 assumed the old state: decode and request again (the stale copies
            // must have landed before the same buffer is written again)
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            pn = decode(wn);
            request(pn, pren, (s + 1) & 1);
        }

        if (s >= burn && lead) {
            if (A.state_trace && c < A.trace_n) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                A.traj_E[o] = E;
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
                int p = 0;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = isg[q];
                    if (q == QP) {
#pragma unroll
                        for (int k = 0; k < BCP_PF_NGRID; ++k) A.traj_idx[((size_t)snap_idx * T.P + p++) * n + c] = ipf[k];
                    }
                }
            }
        }
        if (s >= burn && s == next_snap) {
            ++snap_idx;
            const long long ns = (long long)next_snap + A.traj_every;
            next_snap = ns < s_end ? (int)ns : -1;
        }
    };

    Philox4 wA = bcast(gcur, 0), wB;
    Pf2Prop pA = decode(wA), pB;
    Pre xA, xB;
    request(pA, xA, 0);
    for (int s = 0; s < s_end; s += 2) {
        step(s, wA, pA, xA, wB, pB, xB);
        if (s + 1 < s_end) step(s + 1, wB, pB, xB, wA, pA, xA);
    }

    flush(g_state_counts + state, since_state, s_end);
#pragma unroll
    for (int q = 0; q < R; ++q) {
        flush(g_param_counts + T.r[q].hist_sigma + isg[q], since_sg[q], s_end);
        if (lead) {
            C.isg[(size_t)q * n + c] = isg[q];
            C.sep[(size_t)q * n + c] += acc_r[q];
        }
    }
#pragma unroll
    for (int k = 0; k < BCP_PF_NGRID; ++k) {
        flush(g_param_counts + pf.hist[k] + ipf[k], since_pf[k], s_end);
        if (lead) C.ipf[(size_t)k * n + c] = ipf[k];
    }
    if (lead) {
        C.state[c] = state;
        C.E[c] = E;
        C.accepted[c] += acc_total;
        C.sep[(size_t)R * n + c] += acc_state;
    }
}

template <int R, int NJ>
static cudaError_t launch_pf2_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_pf2_kernel<R, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.sig_bytes + 512 * NJ + 4 * 4 * 256 * NJ);
    if (e != cudaSuccess) return e;
    const int block = 128;                                   // 4 chains per CTA
    const long long threads = C.n_chains * 32;
    const int grid = (int)((threads + block - 1) / block);
    mcmc_pf2_kernel<R, NJ><<<grid, block, T.sig_bytes + 512 * NJ + 4 * 4 * 256 * NJ, st>>>(T, C, A);
    return cudaGetLastError();
}

// cudaErrorNotSupported: the ensemble is outside this kernel's shape (gamma restraints next to the PF one,
// more than 4 restraints or more than 128 residues) -> sampler_pf.cu
cudaError_t launch_mcmc_pf2(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st) {
    if (T.pf.q < 0 || T.R > 4 || T.pf.J > 128 || T.pf.J < 1) return cudaErrorNotSupported;
    for (int q = 0; q < T.R; ++q)
        if (T.r[q].has_gamma) return cudaErrorNotSupported;
    const int nj = (T.pf.J + 31) / 32;
#define BCP_PF2(r)                                                              \
    case r:                                                                     \
        switch (nj) {                                                           \
            case 1: return launch_pf2_one<r, 1>(T, C, A, st);                   \
            case 2: return launch_pf2_one<r, 2>(T, C, A, st);                   \
            case 3: return launch_pf2_one<r, 3>(T, C, A, st);                   \
            default: return launch_pf2_one<r, 4>(T, C, A, st);                  \
        }
    switch (T.R) {
        BCP_PF2(1) BCP_PF2(2) BCP_PF2(3) BCP_PF2(4)
        default: return cudaErrorNotSupported;
    }
#undef BCP_PF2
}

}  // namespace bcp
