// K5: many independent Metropolis chains over (state, sigma index, gamma index).
//
// Replaces PosteriorSampler.sample / neglogP and Restraint_*.compute_neglogP
// (/root/reference/biceps/PosteriorSampler.py:233-374, Restraint.py:356-383, 457-477,
// 572-592).  One thread = one chain, one warp = a batch of 32 chains.  Per step a chain
//   1. generates its Philox4x32-10 block for (chain id, step)          [philox.cuh]
//   2. proposes either a uniform random state or a +-1 move of all parameters of one
//      restraint (periodic wrap)                                       [:298-307]
//   3. re-evaluates only the restraint terms the move touches, in the reference's exact
//      fp64 order: t = ((nlsig[s] + sse/two_sig2[s]) + cnorm) - refsum,
//      E = (energy_i + logZ) + t_0 + t_1 + ...                         [:233-248]
//   4. Metropolis test                                                  [:322-326]
//   5. bookkeeping: acceptance counters, state / parameter histograms (run-length
//      compressed: a counter is flushed only when the state or index changes), thinned
//      snapshots and the optional state trace                          [:329-358]
// The sigma-dependent vectors of all restraints (a few KB) are staged once per CTA into
// shared memory with a 1-D TMA bulk copy; the state-dependent words are gathered from L2.
#include "common.cuh"
#include "exp_det.cuh"
#include "philox.cuh"
#include "tables.cuh"
#include "sampler_common.cuh"

#include <stdlib.h>
#include <vector>

namespace bcp {

template <int R>
__device__ __forceinline__ double restraint_term(const SamplerTables& T, const double* s_sig, int q,
                                                 int state, int sg, int gm) {
    const RestraintDev& rd = T.r[q];
    const double sse = __ldg(rd.sse + (size_t)state * rd.n_gamma + gm);
    const double a = s_sig[rd.sig_off + sg];
    const double d = s_sig[rd.sig_off + rd.n_sigma + sg];
    double t = __dadd_rn(a, __ddiv_rn(sse, d));
    t = __dadd_rn(t, rd.cnorm);
    if (rd.refsum) t = __dsub_rn(t, __ldg(rd.refsum + state));
    return t;
}

template <int R>
__global__ void __launch_bounds__(256)
mcmc_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
            const __grid_constant__ LaunchArgs A) {
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
    const double* energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;

    int state = C.state[c];
    double E = C.E[c];
    int isg[R], igm[R];
    double t[R];
    long long since_sg[R], since_gm[R];
    unsigned int acc_r[R];
    unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        isg[q] = C.isg[(size_t)q * n + c];
        igm[q] = C.igm[(size_t)q * n + c];
        t[q] = restraint_term<R>(T, s_sig, q, state, isg[q], igm[q]);
        since_sg[q] = A.s_begin;
        since_gm[q] = A.s_begin;
        acc_r[q] = 0;
    }
    double Cst = __dadd_rn(__ldg(energy + state), logZ);
    long long since_state = A.s_begin;

    long long next_snap = -1, snap_idx = 0;
    if (A.traj_every > 0) {
        // first snapshot step >= s_begin on the lattice burn + m * traj_every
        long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        next_snap = A.burn + m * A.traj_every;
    }

    for (long long s = A.s_begin; s < A.s_end; ++s) {
        const Philox4 w = step_block(C.seed, chain_id, A.step0 + (unsigned long long)s);
        const bool nuis = w.w0 < T.nuis_thresh;
        const uint64_t pr = (uint64_t)w.w1 * (uint32_t)(nuis ? R : T.S);
        const int hi = (int)(pr >> 32);
        uint32_t f = (uint32_t)pr;
        const int pick = nuis ? hi : -1;
        const int nstate = nuis ? state : hi;

        // proposal: indices and the terms that change
        int nsg = 0, ngm = 0;
        double tn[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            tn[q] = t[q];
            if (!nuis) {
                tn[q] = restraint_term<R>(T, s_sig, q, nstate, isg[q], igm[q]);
            } else if (q == pick) {
                uint64_t z = (uint64_t)f * 3u;
                int v = isg[q] + (int)(z >> 32) - 1;
                f = (uint32_t)z;
                const int ns = T.r[q].n_sigma;
                v = v < 0 ? v + ns : (v >= ns ? v - ns : v);
                nsg = v;
                ngm = igm[q];
                if (T.r[q].has_gamma) {
                    const int ng = T.r[q].n_gamma;
                    z = (uint64_t)f * 3u;
                    int g2 = igm[q] + (int)(z >> 32) - 1;
                    f = (uint32_t)z;
                    g2 = g2 < 0 ? g2 + ng : (g2 >= ng ? g2 - ng : g2);
                    ngm = g2;
                }
                tn[q] = restraint_term<R>(T, s_sig, q, state, nsg, ngm);
            }
        }
        const double Cn = nuis ? Cst : __dadd_rn(__ldg(energy + nstate), logZ);
        double En = Cn;
#pragma unroll
        for (int q = 0; q < R; ++q) En = __dadd_rn(En, tn[q]);

        bool acc;
        if (En < E) {
            acc = true;
        } else {
            acc = metropolis_accept(__dsub_rn(E, En), u2_from_words(w.w2, w.w3));
        }

        if (acc) {
            E = En;
            if (!nuis) {
                if (nstate != state) {
                    const long long cnt = recorded(since_state, s, A.burn);
                    if (cnt) atomicAdd(g_state_counts + state, (unsigned long long)cnt);
                    since_state = s;
                    state = nstate;
                }
                Cst = Cn;
                ++acc_state;
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                t[q] = tn[q];
                if (nuis && q == pick) {
                    ++acc_r[q];
                    if (nsg != isg[q]) {
                        const long long cnt = recorded(since_sg[q], s, A.burn);
                        if (cnt) atomicAdd(g_param_counts + T.r[q].hist_sigma + isg[q], (unsigned long long)cnt);
                        since_sg[q] = s;
                        isg[q] = nsg;
                    }
                    if (ngm != igm[q]) {
                        const long long cnt = recorded(since_gm[q], s, A.burn);
                        if (cnt) atomicAdd(g_param_counts + T.r[q].hist_gamma + igm[q], (unsigned long long)cnt);
                        since_gm[q] = s;
                        igm[q] = ngm;
                    }
                }
            }
            ++acc_total;
        }

        if (s >= A.burn) {
            if (A.state_trace && c < A.trace_n) A.state_trace[(size_t)(s - A.burn) * A.trace_n + c] = state;
            if (s == next_snap) {
                const size_t o = (size_t)snap_idx * n + c;
                A.traj_E[o] = E;
                A.traj_state[o] = state;
                A.traj_accept[o] = acc ? 1 : 0;
                int p = 0;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    A.traj_idx[((size_t)snap_idx * T.P + p) * n + c] = isg[q];
                    ++p;
                    if (T.r[q].has_gamma) {
                        A.traj_idx[((size_t)snap_idx * T.P + p) * n + c] = igm[q];
                        ++p;
                    }
                }
                ++snap_idx;
                next_snap += A.traj_every;
            }
        }
    }

    // flush run-length counters and persist the chain
    {
        const long long cnt = recorded(since_state, A.s_end, A.burn);
        if (cnt) atomicAdd(g_state_counts + state, (unsigned long long)cnt);
    }
    C.state[c] = state;
    C.E[c] = E;
    C.accepted[c] += acc_total;
    C.sep[(size_t)R * n + c] += acc_state;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        long long cnt = recorded(since_sg[q], A.s_end, A.burn);
        if (cnt) atomicAdd(g_param_counts + T.r[q].hist_sigma + isg[q], (unsigned long long)cnt);
        if (T.r[q].has_gamma) {
            cnt = recorded(since_gm[q], A.s_end, A.burn);
            if (cnt) atomicAdd(g_param_counts + T.r[q].hist_gamma + igm[q], (unsigned long long)cnt);
        }
        C.isg[(size_t)q * n + c] = isg[q];
        C.igm[(size_t)q * n + c] = igm[q];
        C.sep[(size_t)q * n + c] += acc_r[q];
    }
}

// energy[K][S] -> energy_t[S][K]
__global__ void transpose_energy_kernel(const double* __restrict__ e, int K, int S, double* __restrict__ et) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [S][K]
    if (i >= (long long)K * S) return;
    const int st = (int)(i / K), l = (int)(i - (long long)st * K);
    et[i] = e[(size_t)l * S + st];
}

// initial state of every chain from the Philox block at step 2^64-1 (oracle/philox.py)
__global__ void init_chains_kernel(long long n, unsigned long long seed, unsigned long long chain_id0,
                                   int S, int R, const int* init_state, int* state, double* E,
                                   int* isg, int* igm, int* ipf, SamplerTables T) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int s0;
    if (init_state) {
        s0 = init_state[c];
    } else {
        const Philox4 w = step_block(seed, chain_id0 + (unsigned long long)c, BCP_INIT_STEP);
        s0 = (int)(((uint64_t)w.w1 * (uint32_t)S) >> 32);
    }
    state[c] = s0;
    E[c] = 1.0e99;
    for (int q = 0; q < R; ++q) {
        isg[(size_t)q * n + c] = T.r[q].n_sigma / 2;
        igm[(size_t)q * n + c] = T.r[q].has_gamma ? T.r[q].n_gamma / 2 : 0;
    }
    if (ipf)
        for (int k = 0; k < BCP_PF_NGRID; ++k) ipf[(size_t)k * n + c] = T.pf.n[k] / 2;
}

// -log P of explicit configurations (PosteriorSampler.neglogP, PosteriorSampler.py:233-248)
__global__ void neglogp_kernel(const SamplerTables T, long long n, const int* lam, const int* state,
                               const int* indices, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = state[i];
    const int l = lam ? lam[i] : 0;
    double E = __dadd_rn(T.energy[(size_t)l * T.S + st], T.logZ[l]);
    int p = 0;
    for (int q = 0; q < T.R; ++q) {
        const RestraintDev& rd = T.r[q];
        const int sg = indices[(size_t)i * T.P + p++];
        const int gm = rd.has_gamma ? indices[(size_t)i * T.P + p++] : 0;
        const double sse = rd.sse[(size_t)st * rd.n_gamma + gm];
        const double* sig = T.sig_tab + rd.sig_off;
        double t = __dadd_rn(sig[sg], __ddiv_rn(sse, sig[rd.n_sigma + sg]));
        t = __dadd_rn(t, rd.cnorm);
        if (rd.refsum) t = __dsub_rn(t, rd.refsum[st]);
        E = __dadd_rn(E, t);
    }
    out[i] = E;
}

// canonical layout (tables.cuh) from the uploaded per-restraint tables: packed records, the gamma
// restraint's rows with periodic halos of 2 and 4, and the range check behind SamplerTables::div_safe
__global__ void build_layouts_kernel(const SamplerTables T, double* rec, double* halo2, double* halo4, int* bad) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int R = T.R;
    const int width = T.gh4_stride > 2 * R ? T.gh4_stride : 2 * R;
    const long long i = idx / width;
    const int k = (int)(idx - i * width);
    if (i >= T.S) return;
    bool flag = false;
    if (k < 2 * R) {
        const int q = k >> 1;
        const RestraintDev& rd = T.r[q];
        double v;
        if (k & 1) v = rd.refsum ? rd.refsum[i] : 0.0;
        else {
            v = rd.has_gamma ? 0.0 : rd.sse[i];
            flag = !(v == 0.0 || (v > 1e-140 && v < 1e140));
        }
        rec[i * (2 * R) + k] = v;
    }
    if (halo4 && k < T.gh4_stride) {
        const int G = T.r[R - 1].n_gamma;
        const double* src = T.r[R - 1].sse + i * G;
        const double v = src[((k - 4) % G + G) % G];
        halo4[i * T.gh4_stride + k] = v;
        if (k < G + 4) halo2[i * (G + 4) + k] = src[((k - 2) % G + G) % G];
        flag = flag || !(v == 0.0 || (v > 1e-140 && v < 1e140));
    }
    if (flag) *bad = 1;
}

// lower bound of -log P of every state over all nuisance parameters (sampler_tab.cu rejects state-move proposals with it):
// per restraint min over the sigma grid of Ndof ln sigma + m / 2 sigma^2 with m = min over gamma of the sse, plus the
// state-dependent constants, minus a safety of 1e-9 (1 + sum of the magnitudes) -- rounding of either side is ~1e-16 relative
__global__ void build_lb_kernel(const SamplerTables T, double* lb) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)T.K * T.S) return;
    const int l = (int)(idx / T.S);
    const long long i = idx - (long long)l * T.S;
    double sum = T.energy[idx] + T.logZ[l];
    double mag = fabs(T.energy[idx]) + fabs(T.logZ[l]);
    for (int q = 0; q < T.R; ++q) {
        const RestraintDev& rd = T.r[q];
        double m = rd.sse[i * rd.n_gamma];
        for (int g = 1; g < rd.n_gamma; ++g) m = fmin(m, rd.sse[i * rd.n_gamma + g]);
        const double* sig = T.sig_tab + rd.sig_off;
        double best = sig[0] + m / sig[rd.n_sigma];
        for (int k = 1; k < rd.n_sigma; ++k) best = fmin(best, sig[k] + m / sig[rd.n_sigma + k]);
        const double ref = rd.refsum ? rd.refsum[i] : 0.0;
        sum += best + rd.cnorm - ref;
        mag += fabs(best) + fabs(rd.cnorm) + fabs(ref);
    }
    lb[idx] = sum - 1e-9 * (1.0 + mag);
}

// K6 on the device-resident snapshots of a sample() call (bcp_chains_score): snapshot m of chain c is
// sample pos0[c] + m of its lambda run (the chains of a run are concatenated, kln_to_kn order).  The
// restraint terms do not depend on lambda: evaluated once, K energies written in
// PosteriorSampler.neglogP order.  m is the fastest thread index, so the K * 8-byte writes coalesce.
__global__ void __launch_bounds__(256)
ukn_traj_kernel(const SamplerTables T, const double* __restrict__ energy_t, long long n_chains, long long n_snap, long long N,
                const long long* __restrict__ pos0, const int* __restrict__ traj_state, const int* __restrict__ traj_idx,
                double* __restrict__ u, int* __restrict__ state_n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_chains * n_snap) return;
    const long long c = idx / n_snap, m = idx - c * n_snap;
    const long long nn = pos0[c] + m;
    const int st = traj_state[m * n_chains + c];
    double t[BCP_MAX_RESTRAINTS];
    int p = 0;
#pragma unroll
    for (int q = 0; q < BCP_MAX_RESTRAINTS; ++q) {
        if (q < T.R) {
            const RestraintDev& rd = T.r[q];
            const int sg = traj_idx[(m * T.P + p++) * n_chains + c];
            const int gm = rd.has_gamma ? traj_idx[(m * T.P + p++) * n_chains + c] : 0;
            const double* sig = T.sig_tab + rd.sig_off;
            double v = __dadd_rn(sig[sg], __ddiv_rn(rd.sse[(size_t)st * rd.n_gamma + gm], sig[rd.n_sigma + sg]));
            v = __dadd_rn(v, rd.cnorm);
            if (rd.refsum) v = __dsub_rn(v, rd.refsum[st]);
            t[q] = v;
        }
    }
    const double* __restrict__ erow = energy_t + (size_t)st * T.K;         // the K energies of the state, contiguous
    for (int l = 0; l < T.K; ++l) {
        double E = __dadd_rn(erow[l], T.logZ[l]);
#pragma unroll
        for (int q = 0; q < BCP_MAX_RESTRAINTS; ++q)
            if (q < T.R) E = __dadd_rn(E, t[q]);
        u[(size_t)l * N + nn] = E;
    }
    state_n[nn] = st;
}

}  // namespace bcp

using namespace bcp;

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct bcp_handle {
    int device = 0;
    int refs = 1;            // the owner + one per live bcp_chains (destruction order is free)
    SamplerTables T{};
    std::vector<void*> allocs;
    int32_t rest_index[BCP_MAX_PARAMS];
    int32_t grid_len[BCP_MAX_PARAMS];
    int32_t p_slot[BCP_MAX_PARAMS];   // where parameter p lives: 0 sigma, 1 gamma, 2 + k PF grid k
    double* d_energy_t = nullptr;     // energy[K][S] transposed to [S][K] for the rescoring kernels (K6)
    cudaStream_t stream = nullptr;    // the handle's own stream: table uploads, layout kernels, bcp_neglogp
    // persistent scratch of bcp_neglogp (the reference's Analysis calls neglogP K^2 x nsnaps times, one configuration
    // per call: Analysis.py:182): device + page-locked host staging for up to NL_CAP configurations
    static constexpr int NL_CAP = 4096;
    unsigned char* nl_dev = nullptr;
    unsigned char* nl_host = nullptr;
};

struct bcp_chains {
    bcp_handle* h = nullptr;
    ChainBuffers C{};
    int n_groups = 0;
    unsigned long long steps_done = 0;   // Philox step counter == reference's `total`
    int* d_lam = nullptr;
    int* d_group = nullptr;
    // outputs of the last sample() call
    long long n_snap = 0, n_trace = 0, trace_n = 0;
    double* d_traj_E = nullptr;
    int* d_traj_state = nullptr;
    unsigned char* d_traj_accept = nullptr;
    int* d_traj_idx = nullptr;
    int* d_trace = nullptr;
    size_t cap_snap = 0, cap_trace = 0;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_chunk = nullptr;
    std::vector<int> h_lam;                 // lambda index of every chain (bcp_chains_score)
    int spec_verdict = 0;                   // 0 undecided, 1 keep the speculative kernel, -1 its batches roll back too often
    int tab_verdict = 0, tab_trials = 0;    // table-driven kernel (its rollback unit is 8 x larger): 0 undecided, 1 keep, -1 not for this ensemble
    double tab_ms_step = 0.0, alt_ms_step = 0.0;   // measured time per step of the last trial launch of the table-driven kernel / of the other choice
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaStream_t copy_stream = nullptr;     // device -> host copies of finished snapshots (bcp_sample)
    long long snaps_piped = 0;
    int last_launches = 0;
    const char* last_kernel = "none";      // the Metropolis kernel the last launch of the last sample call ran
    bool sampled = false;
    // reference-stream mode (bcp_chain_cfg.rng_mode == 1): numpy's MT19937 outputs, one chain
    int rng_mode = 0;
    uint32_t* d_mt_raw = nullptr;
    long long mt_len = 0;
    long long* d_mt_pos = nullptr;
    int* d_mt_dry = nullptr;
};

template <typename T>
static int upload(bcp_handle* h, const T* host, size_t n, const T** dev) {
    T* d = nullptr;
    if (n == 0) { *dev = nullptr; return BCP_OK; }
    BCP_CUDA(pool_alloc(&d, n, h->stream));
    h->allocs.push_back(d);
    // asynchronous on the handle's stream: page-locked caller tables (bcp_host_alloc) stream back to back at the
    // PCIe rate, pageable ones are staged by the driver before the call returns; bcp_create synchronises once
    BCP_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dev = d;
    return BCP_OK;
}

extern "C" int bcp_create(const bcp_tables* t, int device, bcp_handle** out) {
    BCP_REQUIRE(t && out, "bcp_create: NULL argument");
    *out = nullptr;
    BCP_REQUIRE(t->n_states > 0 && t->n_lambda > 0, "bcp_create: n_states and n_lambda must be > 0");
    BCP_REQUIRE(t->n_restraints > 0 && t->n_restraints <= MAXR,
                "bcp_create: n_restraints must be in 1..%d", MAXR);
    BCP_REQUIRE(t->energy && t->logZ && t->restraints, "bcp_create: NULL table pointer");
    int ndev = 0;
    BCP_CUDA(cudaGetDeviceCount(&ndev));
    BCP_REQUIRE(device >= 0 && device < ndev, "bcp_create: device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);

    bcp_handle* h = new bcp_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("bcp_create: cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return BCP_ERR_CUDA;
    }
    SamplerTables& T = h->T;
    T.R = t->n_restraints; T.S = t->n_states; T.K = t->n_lambda;
    T.nuis_thresh = (uint32_t)(((uint64_t)T.R << 32) / (uint64_t)(T.R + 1));
    T.pf.q = -1;
    int rc = BCP_OK;
    auto fail = [&](int code) { bcp_destroy(h); return code; };

    // sigma-dependent vectors of all restraints, contiguous, padded to 16 B for the bulk copy
    std::vector<double> sig;
    int P = 0, HB = 0;
    for (int q = 0; q < T.R; ++q) {
        const bcp_restraint& r = t->restraints[q];
        const bool is_pf = r.kind == BCP_KIND_PFGRID;
        if (!(r.kind == BCP_KIND_SCALAR || r.kind == BCP_KIND_GAMMA || is_pf) || r.n_sigma <= 0 || !r.nlsig ||
            !r.two_sig2 || (!is_pf && !r.sse) || (r.kind == BCP_KIND_GAMMA && r.n_gamma <= 0)) {
            set_error("bcp_create: restraint %d: bad kind/size/pointer", q);
            return fail(BCP_ERR_INVALID);
        }
        if (is_pf) {
            bool ok = T.pf.q < 0 && r.pf_n_obs > 0 && r.pf_grid && r.pf_Nc && r.pf_Nh && r.pf_exp && r.pf_weight &&
                      (r.pf_ref_kind == 0 || r.pf_ref_kind == 1) && P + 1 + BCP_PF_NGRID <= BCP_MAX_PARAMS;
            for (int k = 0; k < BCP_PF_NGRID; ++k) ok = ok && r.pf_n[k] > 0;
            if (!ok) {
                set_error("bcp_create: restraint %d: bad PFGRID description (one PFGRID restraint per ensemble, "
                          "all six grids non-empty, ref uniform or exponential)", q);
                return fail(BCP_ERR_INVALID);
            }
        }
        RestraintDev& rd = T.r[q];
        rd.n_sigma = r.n_sigma;
        rd.n_gamma = r.kind == BCP_KIND_GAMMA ? r.n_gamma : 1;
        rd.cnorm = r.cnorm;
        rd.sig_off = (int)sig.size();
        sig.insert(sig.end(), r.nlsig, r.nlsig + r.n_sigma);
        sig.insert(sig.end(), r.two_sig2, r.two_sig2 + r.n_sigma);
        rd.hist_sigma = HB; HB += rd.n_sigma;
        h->rest_index[P] = q; h->grid_len[P] = rd.n_sigma; h->p_slot[P] = 0; ++P;
        rd.has_gamma = r.kind == BCP_KIND_GAMMA;
        if (rd.has_gamma) {
            rd.hist_gamma = HB; HB += rd.n_gamma;
            h->rest_index[P] = q; h->grid_len[P] = rd.n_gamma; h->p_slot[P] = 1; ++P;
        }
        if (is_pf) {
            PfDev& pf = T.pf;
            pf.q = q; pf.J = r.pf_n_obs; pf.ref_kind = r.pf_ref_kind;
            size_t ngrid = 0, cells = 1;
            for (int k = 0; k < BCP_PF_NGRID; ++k) {
                pf.n[k] = r.pf_n[k]; pf.hist[k] = HB; HB += r.pf_n[k];
                h->rest_index[P] = q; h->grid_len[P] = r.pf_n[k]; h->p_slot[P] = 2 + k; ++P;
                ngrid += r.pf_n[k]; cells *= (size_t)r.pf_n[k];
            }
            const size_t J = (size_t)pf.J;
            if ((rc = upload(h, r.pf_grid, ngrid, &pf.grid)) != BCP_OK) return fail(rc);
            if ((rc = upload(h, r.pf_Nc, (size_t)T.S * pf.n[3] * pf.n[5] * J, &pf.Nc)) != BCP_OK) return fail(rc);
            if ((rc = upload(h, r.pf_Nh, (size_t)T.S * pf.n[4] * pf.n[5] * J, &pf.Nh)) != BCP_OK) return fail(rc);
            if ((rc = upload(h, r.pf_exp, J, &pf.exp)) != BCP_OK) return fail(rc);
            if ((rc = upload(h, r.pf_weight, J, &pf.weight)) != BCP_OK) return fail(rc);
            pf.prior = nullptr;
            if (r.pf_prior && (rc = upload(h, r.pf_prior, cells, &pf.prior)) != BCP_OK) return fail(rc);
            rd.sse = nullptr; rd.refsum = nullptr;
            continue;
        }
        if ((rc = upload(h, r.sse, (size_t)T.S * rd.n_gamma, &rd.sse)) != BCP_OK) return fail(rc);
        rd.refsum = nullptr;
        if (r.has_ref) {
            if (!r.refsum) { set_error("bcp_create: restraint %d: has_ref but refsum is NULL", q); return fail(BCP_ERR_INVALID); }
            if ((rc = upload(h, r.refsum, (size_t)T.S, &rd.refsum)) != BCP_OK) return fail(rc);
        }
    }
    while (sig.size() % 2) sig.push_back(0.0);
    T.P = P; T.HB = HB;
    T.sig_bytes = (int)(sig.size() * sizeof(double));
    if (T.sig_bytes > 200 * 1024) {
        set_error("bcp_create: sigma grids need %d B of shared memory (limit 200 KB)", T.sig_bytes);
        return fail(BCP_ERR_INVALID);
    }
    if ((rc = upload(h, sig.data(), sig.size(), &T.sig_tab)) != BCP_OK) return fail(rc);
    // packed per-state records for the specialised kernel (tables.cuh)
    T.canonical = T.pf.q < 0 ? 1 : 0;
    for (int q = 0; q < T.R; ++q)
        if (T.r[q].has_gamma && (q != T.R - 1 || T.r[q].n_gamma < 2)) T.canonical = 0;
    T.rec_stride = 2 * T.R;
    T.rec = nullptr;
    if (T.canonical) {
        // derived layouts are built on the device from the uploaded tables (a host-side build of the
        // 2 x 13 MB halo tables dominated bcp_create)
        double* d_rec = nullptr;
        if (pool_alloc(&d_rec, (size_t)T.S * T.rec_stride, h->stream) != cudaSuccess) { set_error("bcp_create: out of device memory"); return fail(BCP_ERR_CUDA); }
        h->allocs.push_back(d_rec);
        T.rec = d_rec;
        T.gsse_halo = nullptr; T.gsse_halo4 = nullptr; T.gh4_stride = 0;
        double *d_h2 = nullptr, *d_h4 = nullptr;
        int G = 1;
        if (T.r[T.R - 1].has_gamma) {
            G = T.r[T.R - 1].n_gamma;
            T.gh4_stride = (G + 9 + 1) & ~1;
            if (pool_alloc(&d_h2, (size_t)T.S * (G + 4), h->stream) != cudaSuccess || pool_alloc(&d_h4, (size_t)T.S * T.gh4_stride, h->stream) != cudaSuccess) {
                if (d_h2) h->allocs.push_back(d_h2);
                set_error("bcp_create: out of device memory");
                return fail(BCP_ERR_CUDA);
            }
            h->allocs.push_back(d_h2); h->allocs.push_back(d_h4);
            T.gsse_halo = d_h2; T.gsse_halo4 = d_h4;
        }
        int* d_flag = nullptr;
        if (pool_alloc(&d_flag, 1, h->stream) != cudaSuccess) { set_error("bcp_create: out of device memory"); return fail(BCP_ERR_CUDA); }
        h->allocs.push_back(d_flag);
        cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream);
        const long long work = (long long)T.S * (T.gh4_stride > 2 * T.R ? T.gh4_stride : 2 * T.R);
        build_layouts_kernel<<<(unsigned)((work + 255) / 256), 256, 0, h->stream>>>(T, d_rec, d_h2, d_h4, d_flag);
        BCP_LAUNCHED(1);
        // packed sigma entries for the speculative kernel; 1 / (2 sigma^2) is the correctly rounded
        // IEEE reciprocal (host division), which is what makes the device-side Markstein division exact
        std::vector<double> sig4;
        T.div_safe = 1;
        for (int q = 0; q < T.R; ++q) {
            const bcp_restraint& r = t->restraints[q];
            const int ns = r.n_sigma;
            T.ent_off[q] = (int)(sig4.size() / 3);
            for (int e = 0; e < ns + 4; ++e) {
                const int k = ((e - 2) % ns + ns) % ns;
                const double d = r.two_sig2[k];
                if (!(d > 1e-140 && d < 1e140)) T.div_safe = 0;
                sig4.push_back(r.nlsig[k]); sig4.push_back(d); sig4.push_back(1.0 / d);
            }
        }
        while (sig4.size() % 2) sig4.push_back(0.0);
        T.sig4_bytes = (int)(sig4.size() * sizeof(double));
        if ((rc = upload(h, sig4.data(), sig4.size(), &T.sig4)) != BCP_OK) return fail(rc);
        int flag = 0;
        if (cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            set_error("bcp_create: layout kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(BCP_ERR_CUDA);
        }
        if (flag) T.div_safe = 0;       // some sse value is not finite or far outside the normal range
    }
    if ((rc = upload(h, t->energy, (size_t)T.K * T.S, &T.energy)) != BCP_OK) return fail(rc);
    if ((rc = upload(h, t->logZ, (size_t)T.K, &T.logZ)) != BCP_OK) return fail(rc);
    T.lb = nullptr;
    if (T.canonical && T.div_safe) {
        double* d_lb = nullptr;
        if (pool_alloc(&d_lb, (size_t)T.K * T.S, h->stream) != cudaSuccess) { set_error("bcp_create: out of device memory"); return fail(BCP_ERR_CUDA); }
        h->allocs.push_back(d_lb);
        build_lb_kernel<<<(unsigned)(((long long)T.K * T.S + 127) / 128), 128, 0, h->stream>>>(T, d_lb);
        BCP_LAUNCHED(1);
        T.lb = d_lb;
    }
    if (pool_alloc(&h->d_energy_t, (size_t)T.K * T.S, h->stream) != cudaSuccess) { set_error("bcp_create: out of device memory"); return fail(BCP_ERR_CUDA); }
    h->allocs.push_back(h->d_energy_t);
    transpose_energy_kernel<<<(unsigned)(((long long)T.K * T.S + 255) / 256), 256, 0, h->stream>>>(T.energy, T.K, T.S, h->d_energy_t);
    BCP_LAUNCHED(1);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {       // no caller pointer is kept after the return
        set_error("bcp_create: table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(BCP_ERR_CUDA);
    }
    *out = h;
    return BCP_OK;
}

extern "C" void bcp_destroy(bcp_handle* h) {
    if (!h) return;
    if (--h->refs > 0) return;      // chains still alive: the last bcp_chains_destroy frees the tables
    DeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) pool_free(p);
    pool_free(h->nl_dev);
    if (h->nl_host) cudaFreeHost(h->nl_host);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

namespace bcp {
const SamplerTables& handle_tables(const bcp_handle* h) { return h->T; }
const double* handle_energy_t(const bcp_handle* h) { return h->d_energy_t; }
int handle_device(const bcp_handle* h) { return h->device; }
}  // namespace bcp

extern "C" int bcp_n_params(const bcp_handle* h) { return h ? h->T.P : BCP_ERR_INVALID; }
extern "C" int bcp_param_bins(const bcp_handle* h) { return h ? h->T.HB : BCP_ERR_INVALID; }
extern "C" int bcp_param_layout(const bcp_handle* h, int32_t* rest_index, int32_t* grid_len) {
    BCP_REQUIRE(h, "bcp_param_layout: NULL handle");
    for (int p = 0; p < h->T.P; ++p) {
        if (rest_index) rest_index[p] = h->rest_index[p];
        if (grid_len) grid_len[p] = h->grid_len[p];
    }
    return BCP_OK;
}

extern "C" int bcp_chains_create(bcp_handle* h, const bcp_chain_cfg* cfg, bcp_chains** out) {
    BCP_REQUIRE(h && cfg && out, "bcp_chains_create: NULL argument");
    *out = nullptr;
    BCP_REQUIRE(cfg->n_chains > 0, "bcp_chains_create: n_chains must be > 0");
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    const long long n = cfg->n_chains;
    const int n_groups = cfg->n_groups > 0 ? cfg->n_groups : T.K;
    std::vector<int> lam(n, 0), grp(n, 0);
    for (long long i = 0; i < n; ++i) {
        if (cfg->chain_lambda) lam[i] = cfg->chain_lambda[i];
        BCP_REQUIRE(lam[i] >= 0 && lam[i] < T.K, "bcp_chains_create: chain_lambda[%lld]=%d out of range", i, lam[i]);
        grp[i] = cfg->chain_group ? cfg->chain_group[i] : lam[i];
        BCP_REQUIRE(grp[i] >= 0 && grp[i] < n_groups, "bcp_chains_create: chain_group[%lld]=%d out of range", i, grp[i]);
        if (cfg->init_state)
            BCP_REQUIRE(cfg->init_state[i] >= 0 && cfg->init_state[i] < T.S,
                        "bcp_chains_create: init_state[%lld]=%d out of range", i, cfg->init_state[i]);
    }
    BCP_REQUIRE(cfg->rng_mode == 0 || cfg->rng_mode == 1, "bcp_chains_create: rng_mode must be 0 (Philox) or 1 (MT19937 stream)");
    BCP_REQUIRE(cfg->rng_mode == 0 || n == 1, "bcp_chains_create: the MT19937 stream mode runs one chain");
    bcp_chains* c = new bcp_chains();
    c->h = h;
    c->rng_mode = cfg->rng_mode;
    ++h->refs;
    c->n_groups = n_groups;
    c->h_lam = lam;
    ChainBuffers& C = c->C;
    C.n_chains = n; C.seed = cfg->seed; C.chain_id0 = cfg->chain_id0;
    auto fail = [&](int code) { bcp_chains_destroy(c); return code; };
#define TRY_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(e_)); return fail(BCP_ERR_CUDA); } } while (0)
    TRY_CUDA(pool_alloc(&c->d_lam, n));
    TRY_CUDA(pool_alloc(&c->d_group, n));
    TRY_CUDA(cudaMemcpy(c->d_lam, lam.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    TRY_CUDA(cudaMemcpy(c->d_group, grp.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    C.lam = c->d_lam; C.group = c->d_group;
    TRY_CUDA(pool_alloc(&C.state, n));
    TRY_CUDA(pool_alloc(&C.isg, (size_t)T.R * n));
    TRY_CUDA(pool_alloc(&C.igm, (size_t)T.R * n));
    C.ipf = nullptr;
    if (T.pf.q >= 0) TRY_CUDA(pool_alloc(&C.ipf, (size_t)BCP_PF_NGRID * n));
    TRY_CUDA(pool_alloc(&C.E, n));
    TRY_CUDA(pool_alloc(&C.accepted, n));
    TRY_CUDA(pool_alloc(&C.sep, (size_t)(T.R + 1) * n));
    // both histograms in ONE block, state counts first: a multi-GPU caller all-reduces them with one collective
    TRY_CUDA(pool_alloc(&C.state_counts, (size_t)n_groups * ((size_t)T.S + T.HB)));
    C.param_counts = C.state_counts + (size_t)n_groups * T.S;
    TRY_CUDA(pool_alloc(&C.spec_stats, 8 + 40 * 14 * 4));          // (+ the hand-off trace of a diagnostic build of sampler_tab.cu)
    TRY_CUDA(cudaMemset(C.spec_stats, 0, (8 + 40 * 14 * 4) * sizeof(unsigned long long)));
    TRY_CUDA(cudaMemset(C.accepted, 0, n * sizeof(unsigned long long)));
    TRY_CUDA(cudaMemset(C.sep, 0, (size_t)(T.R + 1) * n * sizeof(unsigned long long)));
    TRY_CUDA(cudaMemset(C.state_counts, 0, (size_t)n_groups * T.S * sizeof(unsigned long long)));
    TRY_CUDA(cudaMemset(C.param_counts, 0, (size_t)n_groups * T.HB * sizeof(unsigned long long)));
    int* d_init = nullptr;
    if (cfg->init_state) {
        TRY_CUDA(pool_alloc(&d_init, n));
        TRY_CUDA(cudaMemcpy(d_init, cfg->init_state, n * sizeof(int), cudaMemcpyHostToDevice));
    }
    const int tpb = 256;
    init_chains_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb>>>(n, C.seed, C.chain_id0, T.S, T.R, d_init,
                                                                 C.state, C.E, C.isg, C.igm, C.ipf, T);
    BCP_LAUNCHED(1);
    TRY_CUDA(cudaGetLastError());
    TRY_CUDA(cudaDeviceSynchronize());
    if (d_init) pool_free(d_init);
    TRY_CUDA(cudaEventCreate(&c->ev0));
    TRY_CUDA(cudaEventCreate(&c->ev1));
#undef TRY_CUDA
    *out = c;
    return BCP_OK;
}

extern "C" void bcp_chains_destroy(bcp_chains* c) {
    if (!c) return;
    DeviceGuard guard(c->h->device);
    if (c->sampled) cudaStreamSynchronize(c->last_stream);      // frees are ordered on the default stream only
    pool_free(c->d_mt_raw); pool_free(c->d_mt_pos); pool_free(c->d_mt_dry);
    pool_free(c->d_lam); pool_free(c->d_group);
    pool_free(c->C.state); pool_free(c->C.isg); pool_free(c->C.igm); pool_free(c->C.ipf); pool_free(c->C.E);
    pool_free(c->C.accepted); pool_free(c->C.sep); pool_free(c->C.state_counts);      // (param_counts lives in the same block)
    pool_free(c->C.spec_stats);
    pool_free(c->d_traj_E); pool_free(c->d_traj_state); pool_free(c->d_traj_accept); pool_free(c->d_traj_idx);
    pool_free(c->d_trace);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_chunk) cudaEventDestroy(c->ev_chunk);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    bcp_handle* h = c->h;
    delete c;
    bcp_destroy(h);                 // drops this block's reference
}

template <int R>
static cudaError_t launch_mcmc(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid,
                               int block, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.sig_bytes);
    if (e != cudaSuccess) return e;
    mcmc_kernel<R><<<grid, block, T.sig_bytes, st>>>(T, C, A);
    return cudaGetLastError();
}

// steps per kernel launch: bounds a launch to a few hundred ms and keeps 32-bit per-launch
// counters far from overflow
static const long long kMaxStepsPerLaunch = 1ll << 22;

// Launches the Metropolis kernel(s) for one sample() call.  pipe != nullptr (host-buffer API with
// page-locked trajectory buffers): the call is cut into a few launches and the snapshots a launch
// completed are copied to the host on a second stream while the next launch runs.
static int sample_launch_impl(bcp_chains* c, const bcp_sample_cfg* cfg, void* stream, const bcp_sample_out* pipe) {
    BCP_REQUIRE(c && cfg, "bcp_sample_launch: NULL argument");
    BCP_REQUIRE(cfg->n_steps > 0 && cfg->burn >= 0 && cfg->traj_every >= 0,
                "bcp_sample_launch: n_steps must be > 0, burn and traj_every >= 0");
    bcp_handle* h = c->h;
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    const long long n = c->C.n_chains;
    cudaStream_t st = (cudaStream_t)stream;

    const long long n_snap = cfg->traj_every > 0 ? (cfg->n_steps + cfg->traj_every - 1) / cfg->traj_every : 0;
    if ((size_t)n_snap * n > c->cap_snap) {
        pool_free(c->d_traj_E); pool_free(c->d_traj_state); pool_free(c->d_traj_accept); pool_free(c->d_traj_idx);
        c->d_traj_E = nullptr; c->d_traj_state = nullptr; c->d_traj_accept = nullptr; c->d_traj_idx = nullptr;
        c->cap_snap = 0;
        BCP_CUDA(pool_alloc(&c->d_traj_E, (size_t)n_snap * n));
        BCP_CUDA(pool_alloc(&c->d_traj_state, (size_t)n_snap * n));
        BCP_CUDA(pool_alloc(&c->d_traj_accept, (size_t)n_snap * n));
        BCP_CUDA(pool_alloc(&c->d_traj_idx, (size_t)n_snap * n * T.P));
        c->cap_snap = (size_t)n_snap * n;
    }
    const long long n_trace = cfg->record_state_trace ? cfg->n_steps : 0;
    const long long trace_n = (cfg->trace_chains > 0 && cfg->trace_chains < n) ? cfg->trace_chains : n;
    if ((size_t)n_trace * trace_n > c->cap_trace) {
        pool_free(c->d_trace); c->d_trace = nullptr; c->cap_trace = 0;
        BCP_CUDA(pool_alloc(&c->d_trace, (size_t)n_trace * trace_n));
        c->cap_trace = (size_t)n_trace * trace_n;
    }
    c->n_snap = n_snap; c->n_trace = n_trace; c->trace_n = trace_n;
    // sep_accepted is per sample() call (PosteriorSampler.py:287)
    BCP_CUDA(cudaMemsetAsync(c->C.sep, 0, (size_t)(T.R + 1) * n * sizeof(unsigned long long), st));

    // grid: spread warps over all SMs first, then grow the CTA
    const long long n_warps = (n + 31) / 32;
    const int sms = num_sms(h->device);
    int wpb = (int)((n_warps + sms - 1) / sms);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
    const int block = wpb * 32;
    const int grid = (int)((n + block - 1) / block);

    LaunchArgs A{};
    A.step0 = c->steps_done;
    A.burn = cfg->burn; A.traj_every = cfg->traj_every; A.n_snap = n_snap;
    A.traj_E = c->d_traj_E; A.traj_state = c->d_traj_state; A.traj_accept = c->d_traj_accept;
    A.traj_idx = c->d_traj_idx; A.state_trace = n_trace ? c->d_trace : nullptr; A.trace_n = trace_n;
    const long long total = cfg->n_steps + cfg->burn;
    // the specialised kernel covers the canonical layout; BCP_FORCE_GENERIC=1 selects the generic one
    // kernel choice: BCP_KERNEL=generic|fast|coop overrides (BCP_FORCE_GENERIC=1 == generic)
    const char* force = getenv("BCP_FORCE_GENERIC");
    const char* kern = getenv("BCP_KERNEL");
    // auto: few chains -> the step time is one warp's instruction stream: spread a chain over L lanes and
    // speculate (spec); many chains -> issue-bound: one thread per chain (fast).  The spec kernel holds one
    // CTA of 128 threads per SM (shared-memory chain blocks) and is flat beyond ~one wave; measured on B200
    // (scripts/probe_sampler.py, probe_c2.py): C3 (L = 4) 1.5e10 steps/s at 4096 chains, 1.7e10 at 32768 vs
    // fast 4.2e9 / 2.9e10; C2 (L = 8) spec 3.7e9 / 7.3e9 at 1024 / 8192 chains vs fast 1.1e9 / 8.5e9.
    const int lanes = T.R <= 2 ? 2 : (T.R <= 4 ? 4 : 8);
    // 8-lane chains (R >= 5): two waves of 8-warp CTAs still beat the thread-per-chain kernel (C2, 8192 chains:
    // 1.18e10 against 1.10e10 chain-steps/s)
    // (4-lane chains: a second wave, 9472 .. 16 384 chains, is only 3 % ahead of the thread-per-chain kernel and a
    // third one far behind -- measured on C3 with scripts/probe_spec_n.py -- so the one-wave threshold stays)
    const long long spec_threads = lanes == 8 ? 2ll * sms * 256 : 2ll * sms * 128;
    // 0 generic, 1 fast, 2 coop, 3 spec, 4 ws (the spec kernel's critical path in consumer warps, Philox / prefetch /
    // bookkeeping in producer warps: sampler_ws.cu; bit-identical, selectable with BCP_KERNEL=ws -- measured on C3 / 4096
    // chains: 6.9 ms per 20 000 steps against 4.6 ms for spec, DESIGN.md section 4, so spec stays the automatic choice)
    int which = !T.canonical ? 0 : (n * lanes <= spec_threads ? 3 : 1);
    // few chains (at most two consumer warps per SM, e.g. the 704 chains of a BICePs-score run): the warp-specialised
    // kernel keeps every consumer warp alone on its SM sub-partition and is 1.2-1.3x faster than spec there
    // (704 chains x 20 000 steps: 3.30 ms against 4.28 ms); BCP_WS_MAX_WARPS_PER_SM overrides the limit (0 = never)
    {
        const char* wl = getenv("BCP_WS_MAX_WARPS_PER_SM");
        const long long lim = wl ? atoll(wl) : 2;
        if (which == 3 && n * lanes <= lim * sms * 32) which = 4;
    }
    // one wave of the table-driven kernel (sampler_tab.cu: 32 chains per CTA, one CTA per SM), long runs: measured on C3 in the
    // steady state (scripts/gpu/r02_tab22.sh, third launch of 100 000 steps) 14.1 ms at 704 .. 4736 chains against 16.5 / 17.6 ms
    // for ws at 704 / 2368 chains (23.4 ms beyond) and 21.2 .. 22.7 ms for spec; BCP_TAB=0 disables the choice.  The kernel
    // has to earn its place per chain block (trial launches below); the kernel chosen here otherwise takes over.
    const int which_no_tab = which;
    {
        const char* tb = getenv("BCP_TAB");
        const char* tw = getenv("BCP_TAB_WAVES");            // (tuning: waves of 32 chains per SM the kernel is considered for)
        // (measured on C3, 50 000 steps: 14.4 / 21.7 / 28.8 ms at two / three / four waves against 32.7 / 32.7 / 49.0 ms for spec and
        // 34-35 ms for the thread-per-chain kernel: four waves still pay)
        const long long waves = tw ? atoll(tw) : 4;
        if ((which == 3 || which == 4 || (which == 1 && waves > 1)) && T.lb && n <= waves * 32ll * sms && !(tb && tb[0] == '0')) which = 5;
    }
    if (kern && T.canonical)
        which = !strcmp(kern, "generic") ? 0 : (!strcmp(kern, "fast") ? 1 : (!strcmp(kern, "spec") ? 3 : (!strcmp(kern, "ws") ? 4 : (!strcmp(kern, "tab") ? 5 : 2))));
    if (force && force[0] == '1') which = 0;
    int fgrid = grid, fblock = block;
    {                                                        // launch shape of the thread-per-chain kernel
        int w = (int)((n_warps + sms - 1) / sms);
        w = w < 1 ? 1 : (w > 4 ? 4 : w);
        fblock = w * 32;
        fgrid = (int)((n + fblock - 1) / fblock);
    }
    if (c->rng_mode == 1) {
        if (!c->d_mt_raw) { set_error("bcp_sample: MT19937 stream mode needs bcp_chains_set_mt_stream first"); return BCP_ERR_STATE; }
        pipe = nullptr;
    }
    BCP_CUDA(cudaEventRecord(c->ev0, st));
    int launches = 0;
    long long chunk = kMaxStepsPerLaunch;
    if (pipe && n_snap >= 8) {
        // ~8 launches (4 for short trajectories), cut on snapshot boundaries: only the last launch's snapshots
        // are copied with nothing to hide behind
        const long long te = cfg->traj_every;
        const long long cuts = n_snap >= 64 ? 8 : 4;
        long long per = ((total + cuts - 1) / cuts + te - 1) / te * te;
        if (per < chunk) chunk = per;
        if (!c->copy_stream) BCP_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        if (!c->ev_chunk) BCP_CUDA(cudaEventCreateWithFlags(&c->ev_chunk, cudaEventDisableTiming));
    } else {
        pipe = nullptr;
    }
    long long snaps_copied = 0;
    // The speculative kernel redoes a whole batch when any chain of a warp accepts a state move: fine when
    // that is rare (C3: 0.2 % of the batches), slower than the thread-per-chain kernel when it is not (a
    // 100-state ensemble with 16 chains per warp: 28 %).  Acceptance is a property of the ensemble, so a
    // long auto-selected run starts with a short pilot launch whose rollback fraction decides; the verdict
    // sticks to the chain block.  Every kernel produces the same bits, so this is invisible in the results.
    const bool spec_auto = (which == 3 || which == 4 || which == 5) && !(kern && T.canonical) && T.pf.q < 0;
    if (spec_auto && c->spec_verdict < 0) which = 1;
    // (a block that is a candidate for the table-driven kernel goes straight to its trial launches: the pilot only follows if
    // those hand the block back)
    const bool tab_first = spec_auto && which == 5 && c->tab_verdict >= 0 && c->spec_verdict == 0 && total >= 2 * 16384;
    const bool tab_kept = spec_auto && which == 5 && c->tab_verdict == 1;     // (its trials are behind it: no pilot either)
    const long long pilot = (spec_auto && c->spec_verdict == 0 && total >= 8192 && !tab_first && !tab_kept) ? 1024 : 0;
    // The table-driven kernel redoes 256 chain-steps where spec redoes 32, and an accepted state move rebuilds a chain's tables:
    // it is kept where its own rollback fraction is small, or where the clock says so.  The first steps of a chain are no measure
    // (indices drifting from the middle of their grids, E = 1e99), so the kernel is tried in launches of 16 384 steps: the first
    // that rolls back less than 6 % of its batches keeps it, one beyond 50 % drops it.  In between a block's first trial is
    // repeated, and from the second on the trial's measured time per step decides, against a timed launch of 4096 steps of the
    // kernel that would run otherwise (every kernel writes the same bits: a trial costs nothing but its own slowness); four lost
    // comparisons hand the block to the other kernels, as do runs shorter than 32 768 steps.
    const bool tab_auto = spec_auto && which == 5;
    const long long tab_trial_steps = 16384;
    if (tab_auto && (c->tab_verdict < 0 || (c->tab_verdict == 0 && total < 2 * tab_trial_steps))) which = which_no_tab;
    if (pilot && which == 5) which = 3;                      // (the pilot itself runs on the spec kernel)
    unsigned long long tab_stats0[2] = {0, 0};
    bool tab_trial = false, alt_trial = false, timing = false;
    long long trial_steps = 0;
    for (long long s0 = 0, s1 = 0; s0 < total; s0 = s1) {
        s1 = (pilot && s0 == 0) ? pilot : s0 + chunk;
        if (alt_trial) {                                     // the timed launch of the other kernel that has just been queued
            float ms = 0.f;
            BCP_CUDA(cudaEventSynchronize(c->ev_t1));
            BCP_CUDA(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
            c->alt_ms_step = (double)ms / (double)trial_steps;
            alt_trial = false;
            if (c->tab_ms_step < c->alt_ms_step) c->tab_verdict = 1;
            else if (++c->tab_trials >= 5) c->tab_verdict = -1;
            which = c->tab_verdict >= 0 && total - s0 >= tab_trial_steps ? 5 : which_no_tab;
            if (getenv("BCP_SPEC_STATS")) fprintf(stderr, "bcp tab trial: %.3f us per step against %.3f us -> verdict %d\n", 1e3 * c->tab_ms_step, 1e3 * c->alt_ms_step, c->tab_verdict);
        }
        if (pilot && s0 == pilot) {
            unsigned long long stats[2] = {0, 0};
            BCP_CUDA(cudaMemcpyAsync(stats, c->C.spec_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
            BCP_CUDA(cudaStreamSynchronize(st));
            const double frac = stats[0] > 0 ? (double)stats[1] / (double)stats[0] : 0.0;
            c->spec_verdict = frac > 0.15 ? -1 : 1;
            which = c->spec_verdict < 0 ? 1 : ((tab_auto && c->tab_verdict >= 0 && total - s0 >= tab_trial_steps) ? 5 : which_no_tab);
            if (getenv("BCP_SPEC_STATS")) fprintf(stderr, "bcp pilot: %llu of %llu spec batches rolled back (%.4f) -> verdict %d\n", stats[1], stats[0], frac, c->spec_verdict);
        }
        if (tab_trial) {                                     // the trial launch that has just been queued: its rollback fraction
            unsigned long long stats[2] = {0, 0};
            BCP_CUDA(cudaMemcpyAsync(stats, c->C.spec_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
            BCP_CUDA(cudaStreamSynchronize(st));
            const double nb = (double)(stats[0] - tab_stats0[0]), nr = (double)(stats[1] - tab_stats0[1]);
            float ms = 0.f;
            BCP_CUDA(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
            c->tab_ms_step = (double)ms / (double)trial_steps;
            tab_trial = false;
            if (nb > 0 && nr < 0.06 * nb) c->tab_verdict = 1;
            else if (nr > 0.5 * nb) c->tab_verdict = -1;                                     // (far off: no second trial)
            else if (c->tab_trials++ == 0) {}                                                // (a block's first trial: its chains' first steps)
            else if (c->alt_ms_step <= 0.0) alt_trial = true;                                // the other kernel's time first
            else if (c->tab_ms_step < c->alt_ms_step) c->tab_verdict = 1;
            else if (++c->tab_trials >= 5) c->tab_verdict = -1;
            if (c->tab_verdict < 0 || alt_trial) which = which_no_tab;
            if (getenv("BCP_SPEC_STATS")) fprintf(stderr, "bcp tab trial: %.0f of %.0f batches rolled back, %.3f us per step -> verdict %d\n", nr, nb, 1e3 * c->tab_ms_step, c->tab_verdict);
        }
        if (alt_trial) {
            if (s1 > s0 + 4096) s1 = s0 + 4096;
            timing = true;
        } else if (tab_auto && which == 5 && c->tab_verdict == 0 && !(pilot && s0 == 0)) {
            BCP_CUDA(cudaMemcpyAsync(tab_stats0, c->C.spec_stats, sizeof(tab_stats0), cudaMemcpyDeviceToHost, st));
            BCP_CUDA(cudaStreamSynchronize(st));
            if (s1 > s0 + tab_trial_steps) s1 = s0 + tab_trial_steps;
            tab_trial = true;
            timing = true;
        }
        if (s1 > total) s1 = total;
        if (timing) {
            if (!c->ev_t0) { BCP_CUDA(cudaEventCreate(&c->ev_t0)); BCP_CUDA(cudaEventCreate(&c->ev_t1)); }
            BCP_CUDA(cudaEventRecord(c->ev_t0, st));
            trial_steps = s1 - s0;
        }
        A.s_begin = s0;
        A.s_end = s1;
        const bool fast = which == 1;
        cudaError_t e;
        const char* kname = "generic";
        if (c->rng_mode == 1) {
            const MtStream M{c->d_mt_raw, c->mt_len, c->d_mt_pos, c->d_mt_dry};
            e = launch_mcmc_mt(T, c->C, A, M, st);
            kname = "mt";
        }
        else if (T.pf.q >= 0) {
            // BCP_PF_KERNEL=generic forces the general warp-per-chain kernel (tests run both)
            const char* pk = getenv("BCP_PF_KERNEL");
            e = (pk && !strcmp(pk, "generic")) ? cudaErrorNotSupported : launch_mcmc_pf2(T, c->C, A, st);
            kname = "pf2";
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_pf(T, c->C, A, st); kname = "pf"; }
        }
        else if (which == 5) {
            e = launch_mcmc_tab(T, c->C, A, st);
            kname = "tab";
            if (e == cudaErrorNotSupported && kern && !strcmp(kern, "tab")) { (void)cudaGetLastError(); e = launch_mcmc_ws(T, c->C, A, st, sms); kname = "ws"; }
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_spec(T, c->C, A, st, sms); kname = "spec"; }
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_coop(T, c->C, A, st, sms); kname = "coop"; }
        }
        else if (which == 4) {
            e = launch_mcmc_ws(T, c->C, A, st, sms);
            kname = "ws";
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_spec(T, c->C, A, st, sms); kname = "spec"; }
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_coop(T, c->C, A, st, sms); kname = "coop"; }
        }
        else if (which == 3) {
            e = launch_mcmc_spec(T, c->C, A, st, sms);
            kname = "spec";
            if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); e = launch_mcmc_coop(T, c->C, A, st, sms); kname = "coop"; }
        }
        else if (which == 2) { e = launch_mcmc_coop(T, c->C, A, st, sms); kname = "coop"; }
        else if (fast) { e = launch_mcmc_fast(T, c->C, A, fgrid, fblock, st); kname = "fast"; }
        else switch (T.R) {
            case 1: e = launch_mcmc<1>(T, c->C, A, grid, block, st); break;
            case 2: e = launch_mcmc<2>(T, c->C, A, grid, block, st); break;
            case 3: e = launch_mcmc<3>(T, c->C, A, grid, block, st); break;
            case 4: e = launch_mcmc<4>(T, c->C, A, grid, block, st); break;
            case 5: e = launch_mcmc<5>(T, c->C, A, grid, block, st); break;
            case 6: e = launch_mcmc<6>(T, c->C, A, grid, block, st); break;
            case 7: e = launch_mcmc<7>(T, c->C, A, grid, block, st); break;
            default: e = launch_mcmc<8>(T, c->C, A, grid, block, st); break;
        }
        if (e != cudaSuccess) { set_error("mcmc_kernel launch failed: %s", cudaGetErrorString(e)); return BCP_ERR_CUDA; }
        c->last_kernel = kname;
        ++launches;
        if (timing) { BCP_CUDA(cudaEventRecord(c->ev_t1, st)); timing = false; }
        if (pipe) {
            // snapshots m with burn + m * traj_every < s_end are final once this launch has finished
            long long done = A.s_end <= cfg->burn ? 0 : (A.s_end - cfg->burn + cfg->traj_every - 1) / cfg->traj_every;
            if (done > n_snap) done = n_snap;
            if (done > snaps_copied) {
                const size_t lo = (size_t)snaps_copied * n, cnt = (size_t)(done - snaps_copied) * n;
                BCP_CUDA(cudaEventRecord(c->ev_chunk, st));
                BCP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_chunk, 0));
                if (pipe->traj_energy) BCP_CUDA(cudaMemcpyAsync(pipe->traj_energy + lo, c->d_traj_E + lo, cnt * 8, cudaMemcpyDeviceToHost, c->copy_stream));
                if (pipe->traj_state) BCP_CUDA(cudaMemcpyAsync(pipe->traj_state + lo, c->d_traj_state + lo, cnt * 4, cudaMemcpyDeviceToHost, c->copy_stream));
                if (pipe->traj_accept) BCP_CUDA(cudaMemcpyAsync(pipe->traj_accept + lo, c->d_traj_accept + lo, cnt, cudaMemcpyDeviceToHost, c->copy_stream));
                if (pipe->traj_indices) BCP_CUDA(cudaMemcpyAsync(pipe->traj_indices + lo * T.P, c->d_traj_idx + lo * T.P, cnt * T.P * 4, cudaMemcpyDeviceToHost, c->copy_stream));
                snaps_copied = done;
            }
        }
    }
    c->snaps_piped = pipe ? snaps_copied : 0;
    BCP_LAUNCHED(launches);
    BCP_CUDA(cudaEventRecord(c->ev1, st));
    c->steps_done += (unsigned long long)total;
    c->last_stream = st;
    c->last_launches = launches;
    c->sampled = true;
    return BCP_OK;
}

extern "C" int bcp_chains_set_mt_stream(bcp_chains* c, const uint32_t* raw, int64_t n_raw) {
    BCP_REQUIRE(c && raw && n_raw > 0, "bcp_chains_set_mt_stream: NULL argument or empty stream");
    if (c->rng_mode != 1) { set_error("bcp_chains_set_mt_stream: the chain block was not created with rng_mode = 1"); return BCP_ERR_STATE; }
    DeviceGuard guard(c->h->device);
    if (c->sampled) BCP_CUDA(cudaStreamSynchronize(c->last_stream));
    if (n_raw > c->mt_len || !c->d_mt_raw) {
        pool_free(c->d_mt_raw); c->d_mt_raw = nullptr;
        BCP_CUDA(pool_alloc(&c->d_mt_raw, (size_t)n_raw));
    }
    if (!c->d_mt_pos) BCP_CUDA(pool_alloc(&c->d_mt_pos, 1));
    if (!c->d_mt_dry) BCP_CUDA(pool_alloc(&c->d_mt_dry, 1));
    c->mt_len = n_raw;
    BCP_CUDA(cudaMemcpy(c->d_mt_raw, raw, (size_t)n_raw * sizeof(uint32_t), cudaMemcpyHostToDevice));
    BCP_CUDA(cudaMemset(c->d_mt_pos, 0, sizeof(long long)));
    BCP_CUDA(cudaMemset(c->d_mt_dry, 0, sizeof(int)));
    return BCP_OK;
}

extern "C" int bcp_chains_mt_consumed(bcp_chains* c, int64_t* n_consumed) {
    BCP_REQUIRE(c && n_consumed, "bcp_chains_mt_consumed: NULL argument");
    if (c->rng_mode != 1 || !c->d_mt_pos) { set_error("bcp_chains_mt_consumed: no MT19937 stream set"); return BCP_ERR_STATE; }
    DeviceGuard guard(c->h->device);
    if (c->sampled) BCP_CUDA(cudaStreamSynchronize(c->last_stream));
    long long pos = 0;
    int dry = 0;
    BCP_CUDA(cudaMemcpy(&pos, c->d_mt_pos, sizeof(pos), cudaMemcpyDeviceToHost));
    BCP_CUDA(cudaMemcpy(&dry, c->d_mt_dry, sizeof(dry), cudaMemcpyDeviceToHost));
    *n_consumed = pos;
    if (dry) { set_error("the MT19937 stream ran out after %lld outputs: hand over a longer one", pos); return BCP_ERR_STATE; }
    return BCP_OK;
}

extern "C" int bcp_sample_launch(bcp_chains* c, const bcp_sample_cfg* cfg, void* stream) {
    return sample_launch_impl(c, cfg, stream, nullptr);
}

extern "C" int bcp_chains_sync(bcp_chains* c) {
    BCP_REQUIRE(c, "bcp_chains_sync: NULL argument");
    DeviceGuard guard(c->h->device);
    BCP_CUDA(cudaStreamSynchronize(c->last_stream));
    if (getenv("BCP_SPEC_STATS")) {       // diagnostics: batches run / rolled back by the speculative kernels so far
        unsigned long long stats[8] = {0};
        BCP_CUDA(cudaMemcpy(stats, c->C.spec_stats, sizeof(stats), cudaMemcpyDeviceToHost));
        fprintf(stderr, "bcp spec stats: %llu batches, %llu rolled back (guard band %llu, window %llu, state move %llu; %llu lower-bound resolutions, %llu accepted)\n",
                stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], stats[6]);
        if (getenv("BCP_TAB_STATS")) {    // (a TAB_STATS build of the table-driven kernel: cycles and counts of lane 0's second looks, by outcome)
            unsigned long long x[6] = {0};
            BCP_CUDA(cudaMemcpy(x, c->C.spec_stats + 7, sizeof(x), cudaMemcpyDeviceToHost));
            fprintf(stderr, "bcp tab second looks: resolved %llu (%llu cycles), replayed %llu (%llu cycles), careful %llu (%llu cycles)\n", x[3], x[0], x[4], x[1], x[5], x[2]);
        }
        if (getenv("BCP_TAB_TRACE")) {
            static unsigned long long tr[40 * 14 * 4];
            BCP_CUDA(cudaMemcpy(tr, c->C.spec_stats + 8, sizeof(tr), cudaMemcpyDeviceToHost));
            for (int b = 0; b < 40; ++b) {
                fprintf(stderr, "trace b=%d", 1000 + b);
                for (int w = 0; w < 14; ++w)
                    for (int e = 0; e < 4; ++e)
                        if (tr[(b * 14 + w) * 4 + e]) fprintf(stderr, " w%d.%d=%llu", w, e, tr[(b * 14 + w) * 4 + e] - tr[0]);
                fprintf(stderr, "\n");
            }
        }
    }
    return BCP_OK;
}

extern "C" const char* bcp_chains_last_kernel(bcp_chains* c) {
    return c ? c->last_kernel : "none";
}

extern "C" int bcp_chains_last_kernel_ms(bcp_chains* c, float* ms, int32_t* n_launches) {
    BCP_REQUIRE(c && ms, "bcp_chains_last_kernel_ms: NULL argument");
    if (!c->sampled) { set_error("bcp_chains_last_kernel_ms: no sample launched yet"); return BCP_ERR_STATE; }
    DeviceGuard guard(c->h->device);
    BCP_CUDA(cudaEventSynchronize(c->ev1));
    BCP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (n_launches) *n_launches = c->last_launches;
    return BCP_OK;
}

extern "C" int bcp_chains_device_hist(bcp_chains* c, int64_t** d_state_counts, int64_t** d_param_counts) {
    BCP_REQUIRE(c, "bcp_chains_device_hist: NULL argument");
    if (d_state_counts) *d_state_counts = (int64_t*)c->C.state_counts;
    if (d_param_counts) *d_param_counts = (int64_t*)c->C.param_counts;
    return BCP_OK;
}

extern "C" int bcp_chains_fetch(bcp_chains* c, bcp_sample_out* o) {
    BCP_REQUIRE(c && o, "bcp_chains_fetch: NULL argument");
    bcp_handle* h = c->h;
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    const size_t n = (size_t)c->C.n_chains;
    BCP_CUDA(cudaStreamSynchronize(c->last_stream));
#define D2H(dst, src, count) do { if (dst) BCP_CUDA(cudaMemcpy(dst, src, (count), cudaMemcpyDeviceToHost)); } while (0)
    D2H(o->state_counts, c->C.state_counts, (size_t)c->n_groups * T.S * 8);
    D2H(o->param_counts, c->C.param_counts, (size_t)c->n_groups * T.HB * 8);
    D2H(o->accepted, c->C.accepted, n * 8);
    if (o->total) for (size_t i = 0; i < n; ++i) o->total[i] = (int64_t)c->steps_done;
    D2H(o->state, c->C.state, n * 4);
    D2H(o->energy, c->C.E, n * 8);
    if (o->indices || o->sep_accepted) {
        std::vector<int> isg((size_t)T.R * n), igm((size_t)T.R * n), ipf;
        std::vector<unsigned long long> sep((size_t)(T.R + 1) * n);
        BCP_CUDA(cudaMemcpy(isg.data(), c->C.isg, isg.size() * 4, cudaMemcpyDeviceToHost));
        BCP_CUDA(cudaMemcpy(igm.data(), c->C.igm, igm.size() * 4, cudaMemcpyDeviceToHost));
        if (c->C.ipf) {
            ipf.resize((size_t)BCP_PF_NGRID * n);
            BCP_CUDA(cudaMemcpy(ipf.data(), c->C.ipf, ipf.size() * 4, cudaMemcpyDeviceToHost));
        }
        BCP_CUDA(cudaMemcpy(sep.data(), c->C.sep, sep.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            for (int p = 0; p < T.P; ++p) {
                const int q = h->rest_index[p], slot = h->p_slot[p];
                if (o->indices)
                    o->indices[i * T.P + p] = slot == 0 ? isg[(size_t)q * n + i]
                                              : (slot == 1 ? igm[(size_t)q * n + i] : ipf[(size_t)(slot - 2) * n + i]);
                if (o->sep_accepted) o->sep_accepted[i * (T.P + 1) + p] = (int64_t)sep[(size_t)q * n + i];
            }
            if (o->sep_accepted) o->sep_accepted[i * (T.P + 1) + T.P] = (int64_t)sep[(size_t)T.R * n + i];
        }
    }
    if (c->sampled) {
        D2H(o->traj_energy, c->d_traj_E, (size_t)c->n_snap * n * 8);
        D2H(o->traj_state, c->d_traj_state, (size_t)c->n_snap * n * 4);
        D2H(o->traj_accept, c->d_traj_accept, (size_t)c->n_snap * n);
        D2H(o->traj_indices, c->d_traj_idx, (size_t)c->n_snap * n * T.P * 4);
        if (o->state_trace) {
            if (!c->n_trace) { set_error("bcp_chains_fetch: state_trace requested but not recorded"); return BCP_ERR_STATE; }
            D2H(o->state_trace, c->d_trace, (size_t)c->n_trace * c->trace_n * 4);
        }
    }
#undef D2H
    return BCP_OK;
}

extern "C" int bcp_chains_set_state(bcp_chains* c, const int32_t* state, const int32_t* indices, const double* energy) {
    BCP_REQUIRE(c, "bcp_chains_set_state: NULL argument");
    bcp_handle* h = c->h;
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    const size_t n = (size_t)c->C.n_chains;
    BCP_CUDA(cudaStreamSynchronize(c->last_stream));
    if (state) {
        for (size_t i = 0; i < n; ++i)
            BCP_REQUIRE(state[i] >= 0 && state[i] < T.S, "bcp_chains_set_state: state[%zu] out of range", i);
        BCP_CUDA(cudaMemcpy(c->C.state, state, n * 4, cudaMemcpyHostToDevice));
    }
    if (energy) BCP_CUDA(cudaMemcpy(c->C.E, energy, n * 8, cudaMemcpyHostToDevice));
    if (indices) {
        std::vector<int> isg((size_t)T.R * n), igm((size_t)T.R * n, 0), ipf((size_t)BCP_PF_NGRID * n, 0);
        for (size_t i = 0; i < n; ++i)
            for (int p = 0; p < T.P; ++p) {
                const int v = indices[i * T.P + p], q = h->rest_index[p], slot = h->p_slot[p];
                BCP_REQUIRE(v >= 0 && v < h->grid_len[p], "bcp_chains_set_state: indices[%zu][%d] out of range", i, p);
                if (slot == 0) isg[(size_t)q * n + i] = v;
                else if (slot == 1) igm[(size_t)q * n + i] = v;
                else ipf[(size_t)(slot - 2) * n + i] = v;
            }
        BCP_CUDA(cudaMemcpy(c->C.isg, isg.data(), isg.size() * 4, cudaMemcpyHostToDevice));
        BCP_CUDA(cudaMemcpy(c->C.igm, igm.data(), igm.size() * 4, cudaMemcpyHostToDevice));
        if (c->C.ipf) BCP_CUDA(cudaMemcpy(c->C.ipf, ipf.data(), ipf.size() * 4, cudaMemcpyHostToDevice));
    }
    return BCP_OK;
}

static bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int bcp_sample(bcp_chains* c, const bcp_sample_cfg* cfg, bcp_sample_out* out) {
    const bool pipe = out && cfg && cfg->traj_every > 0 && out->traj_energy && is_pinned(out->traj_energy) &&
                      is_pinned(out->traj_state) && is_pinned(out->traj_accept) && is_pinned(out->traj_indices);
    int rc = sample_launch_impl(c, cfg, nullptr, pipe ? out : nullptr);
    if (rc != BCP_OK) return rc;
    if (!out) return bcp_chains_sync(c);
    if (c->snaps_piped > 0) {
        // the trajectory went out chunk by chunk on the copy stream; fetch everything else
        DeviceGuard guard(c->h->device);
        bcp_sample_out rest = *out;
        BCP_REQUIRE(c->snaps_piped == c->n_snap, "bcp_sample: internal error, %lld of %lld snapshots copied",
                    c->snaps_piped, c->n_snap);
        rest.traj_energy = nullptr; rest.traj_state = nullptr; rest.traj_accept = nullptr; rest.traj_indices = nullptr;
        rc = bcp_chains_fetch(c, &rest);
        BCP_CUDA(cudaStreamSynchronize(c->copy_stream));
        return rc;
    }
    return bcp_chains_fetch(c, out);
}

// series[(chain * P + p)][t] = grid value of parameter p at snapshot t of the chain (traj_idx is [n_snap][P][n])
__global__ void __launch_bounds__(256)
gather_traces_kernel(const int* __restrict__ traj_idx, const double* __restrict__ grid, const int* __restrict__ off,
                     long long n, int P, long long n_snap, double* __restrict__ series) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [n_snap][P][n], chain fastest
    if (i >= n_snap * P * n) return;
    const long long chain = i % n;
    const int p = (int)((i / n) % P);
    const long long t = i / (n * P);
    series[((size_t)chain * P + p) * n_snap + t] = grid[off[p] + traj_idx[i]];
}

extern "C" int bcp_chains_autocorr(bcp_chains* c, const double* grid_values, int32_t max_tau, int32_t normalize,
                                   double* curves, double* tau) {
    BCP_REQUIRE(c && grid_values && curves, "bcp_chains_autocorr: NULL argument");
    BCP_REQUIRE(c->sampled && c->n_snap >= 2, "bcp_chains_autocorr: needs the snapshots of a sample() call (traj_every > 0)");
    BCP_REQUIRE(max_tau >= 0, "bcp_chains_autocorr: max_tau must be >= 0");
    bcp_handle* h = c->h;
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    const long long n = c->C.n_chains, ns = c->n_snap;
    const int P = T.P;
    cudaStream_t st = c->last_stream;
    std::vector<int> off(P);
    int acc = 0;
    for (int p = 0; p < P; ++p) { off[p] = acc; acc += h->grid_len[p]; }
    double *d_grid = nullptr, *d_series = nullptr, *d_curves = nullptr, *d_tau = nullptr;
    int* d_off = nullptr;
    const size_t n_series = (size_t)n * P, n_out = n_series * ((size_t)max_tau + 1);
    int rc = BCP_OK;
    do {
#define STEP(x) if ((x) != cudaSuccess) { set_error("bcp_chains_autocorr: %s: %s", #x, cudaGetErrorString(cudaGetLastError())); rc = BCP_ERR_CUDA; break; }
        STEP(bcp::pool_malloc((void**)&d_grid, sizeof(double) * (size_t)T.HB, st));
        STEP(bcp::pool_malloc((void**)&d_off, sizeof(int) * (size_t)P, st));
        STEP(bcp::pool_malloc((void**)&d_series, sizeof(double) * n_series * (size_t)ns, st));
        STEP(bcp::pool_malloc((void**)&d_curves, sizeof(double) * n_out, st));
        STEP(bcp::pool_malloc((void**)&d_tau, sizeof(double) * n_series, st));
        STEP(cudaMemcpyAsync(d_grid, grid_values, sizeof(double) * (size_t)T.HB, cudaMemcpyHostToDevice, st));
        STEP(cudaMemcpyAsync(d_off, off.data(), sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, st));
        const long long work = ns * P * n;
        gather_traces_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(c->d_traj_idx, d_grid, d_off, n, P, ns, d_series);
        BCP_LAUNCHED(1);
        STEP(cudaGetLastError());
        rc = autocorr_run(d_series, (int32_t)n_series, ns, max_tau, normalize, st, d_curves, d_tau);
        if (rc != BCP_OK) break;
        STEP(cudaMemcpyAsync(curves, d_curves, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
        if (tau) STEP(cudaMemcpyAsync(tau, d_tau, sizeof(double) * n_series, cudaMemcpyDeviceToHost, st));
        STEP(cudaStreamSynchronize(st));
#undef STEP
    } while (0);
    cudaFreeAsync(d_grid, st); cudaFreeAsync(d_off, st); cudaFreeAsync(d_series, st); cudaFreeAsync(d_curves, st);
    cudaFreeAsync(d_tau, st);
    return rc;
}

// sample order of a chain block's snapshots: runs (lambda) in order, the chains of a run in chain order, a chain's
// snapshots in time order (== Analysis.py:214 kln_to_kn order); pos0[i] = first column of chain i
static long long snapshot_layout(const bcp_chains* c, std::vector<long long>& pos0, std::vector<int64_t>& N_k) {
    const long long n = c->C.n_chains, ns = c->n_snap;
    const int K = c->h->T.K;
    std::vector<long long> cnt(K, 0), off(K + 1, 0), seen(K, 0);
    pos0.assign(n, 0);
    for (long long i = 0; i < n; ++i) ++cnt[c->h_lam[i]];
    for (int k = 0; k < K; ++k) off[k + 1] = off[k] + cnt[k] * ns;
    for (long long i = 0; i < n; ++i) { const int k = c->h_lam[i]; pos0[i] = off[k] + seen[k] * ns; ++seen[k]; }
    N_k.assign(K, 0);
    for (int k = 0; k < K; ++k) N_k[k] = cnt[k] * ns;
    return off[K];
}

static int chains_ukn_launch(bcp_chains* c, const std::vector<long long>& pos0, long long N, double* d_u, int* d_sn,
                             cudaStream_t st) {
    bcp_handle* h = c->h;
    const long long n = c->C.n_chains, ns = c->n_snap;
    long long* d_pos = nullptr;
    BCP_CUDA(pool_alloc(&d_pos, (size_t)n));
    cudaError_t e = cudaMemcpyAsync(d_pos, pos0.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ukn_traj_kernel<<<(unsigned)((n * ns + 255) / 256), 256, 0, st>>>(h->T, h->d_energy_t, n, ns, N, d_pos, c->d_traj_state,
                                                                        c->d_traj_idx, d_u, d_sn);
        BCP_LAUNCHED(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // pos0 is the caller's; d_pos is freed on the default stream
    pool_free(d_pos);
    if (e != cudaSuccess) { set_error("bcp_chains_ukn: %s", cudaGetErrorString(e)); return BCP_ERR_CUDA; }
    return BCP_OK;
}

extern "C" int bcp_chains_n_samples(bcp_chains* c, int64_t* n_samples, int64_t* N_k_out) {
    BCP_REQUIRE(c && n_samples, "bcp_chains_n_samples: NULL argument");
    BCP_REQUIRE(c->sampled && c->n_snap > 0, "bcp_chains_n_samples: no snapshots (sample with traj_every > 0 first)");
    std::vector<long long> pos0;
    std::vector<int64_t> N_k;
    *n_samples = snapshot_layout(c, pos0, N_k);
    if (N_k_out) for (int k = 0; k < c->h->T.K; ++k) N_k_out[k] = N_k[k];
    return BCP_OK;
}

extern "C" int bcp_chains_ukn_dev(bcp_chains* c, double* d_u_kn, int32_t* d_state_n, void* stream) {
    BCP_REQUIRE(c && d_u_kn, "bcp_chains_ukn_dev: NULL argument");
    BCP_REQUIRE(c->sampled && c->n_snap > 0, "bcp_chains_ukn_dev: no snapshots (sample with traj_every > 0 first)");
    BCP_REQUIRE(c->h->T.pf.q < 0, "bcp_chains_ukn_dev: ensembles with a PFGRID restraint go through bcp_ukn");
    DeviceGuard guard(c->h->device);
    std::vector<long long> pos0;
    std::vector<int64_t> N_k;
    const long long N = snapshot_layout(c, pos0, N_k);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != c->last_stream) BCP_CUDA(cudaStreamSynchronize(c->last_stream));     // the snapshots must be complete
    int* d_sn = d_state_n;
    if (!d_sn) BCP_CUDA(pool_alloc(&d_sn, (size_t)N));
    const int rc = chains_ukn_launch(c, pos0, N, d_u_kn, d_sn, st);
    if (!d_state_n) pool_free(d_sn);
    return rc;
}

extern "C" int bcp_chains_score(bcp_chains* c, int32_t want_pops, double tol, int32_t max_iter, int64_t* N_k_out,
                                double* f_k, double* theta, double* pops, double* dpops, int32_t* iters) {
    BCP_REQUIRE(c && f_k && theta, "bcp_chains_score: NULL argument");
    BCP_REQUIRE(c->sampled && c->n_snap > 0, "bcp_chains_score: no snapshots (sample with traj_every > 0 first)");
    bcp_handle* h = c->h;
    const SamplerTables& T = h->T;
    BCP_REQUIRE(T.pf.q < 0, "bcp_chains_score: ensembles with a PFGRID restraint go through bcp_ukn + bcp_mbar");
    DeviceGuard guard(h->device);
    const int K = T.K;
    std::vector<long long> pos0;
    std::vector<int64_t> N_k;
    const long long N = snapshot_layout(c, pos0, N_k);
    if (N_k_out) for (int k = 0; k < K; ++k) N_k_out[k] = N_k[k];
    cudaStream_t st = c->last_stream;
    double* d_u = nullptr; int* d_sn = nullptr;
    int rc = BCP_OK;
    do {
#define STEP(x) if ((x) != cudaSuccess) { set_error("bcp_chains_score: %s: %s", #x, cudaGetErrorString(cudaGetLastError())); rc = BCP_ERR_CUDA; break; }
        STEP(pool_alloc(&d_u, (size_t)K * N));
        STEP(pool_alloc(&d_sn, (size_t)N));
#undef STEP
        rc = chains_ukn_launch(c, pos0, N, d_u, d_sn, st);
        if (rc != BCP_OK) break;
        rc = bcp_mbar_dev(d_u, N_k.data(), K, N, want_pops ? d_sn : nullptr, want_pops ? T.S : 0, h->device, st, tol, max_iter,
                          f_k, theta, pops, dpops, iters);
    } while (0);
    cudaStreamSynchronize(st);
    pool_free(d_u); pool_free(d_sn);
    return rc;
}

extern "C" int bcp_neglogp(bcp_handle* h, int64_t n, const int32_t* lam, const int32_t* state,
                           const int32_t* indices, double* energy_out) {
    BCP_REQUIRE(h && state && indices && energy_out, "bcp_neglogp: NULL argument");
    if (n <= 0) return BCP_OK;
    DeviceGuard guard(h->device);
    const SamplerTables& T = h->T;
    for (int64_t i = 0; i < n; ++i) {
        BCP_REQUIRE(state[i] >= 0 && state[i] < T.S, "bcp_neglogp: state[%lld] out of range", (long long)i);
        BCP_REQUIRE(!lam || (lam[i] >= 0 && lam[i] < T.K), "bcp_neglogp: lam[%lld] out of range", (long long)i);
        for (int p = 0; p < T.P; ++p)
            BCP_REQUIRE(indices[i * T.P + p] >= 0 && indices[i * T.P + p] < h->grid_len[p],
                        "bcp_neglogp: indices[%lld][%d] out of range", (long long)i, p);
    }
    // per configuration: lam (4), state (4), indices (4 P), energy (8); the scratch of the handle is reused from call to
    // call (no allocation, page-locked staging, ONE synchronisation), larger batches go through it in pieces
    cudaStream_t st = h->stream;
    const size_t cap = bcp_handle::NL_CAP;
    const size_t per = 8 + 4 + 4 + 4 * (size_t)T.P;
    if (!h->nl_dev) {
        BCP_CUDA(pool_malloc((void**)&h->nl_dev, cap * per, st));
        BCP_CUDA(cudaHostAlloc((void**)&h->nl_host, cap * per, cudaHostAllocDefault));
    }
    for (int64_t i0 = 0; i0 < n; i0 += (int64_t)cap) {
        const size_t m = (size_t)std::min<int64_t>((int64_t)cap, n - i0);
        // layout (device and host alike): energy[m] | lam[m] | state[m] | indices[m][P]
        double* h_out = (double*)h->nl_host;
        int32_t* h_lam = (int32_t*)(h->nl_host + 8 * cap);
        int32_t* h_state = h_lam + cap;
        int32_t* h_idx = h_state + cap;
        double* d_out = (double*)h->nl_dev;
        int* d_lam = (int*)(h->nl_dev + 8 * cap);
        int* d_state = d_lam + cap;
        int* d_idx = d_state + cap;
        if (lam) memcpy(h_lam, lam + i0, m * 4);
        memcpy(h_state, state + i0, m * 4);
        memcpy(h_idx, indices + i0 * T.P, m * 4 * (size_t)T.P);
        // lam | state | indices are contiguous when m == cap; copy the three pieces (a single configuration: 3 small copies)
        if (lam) BCP_CUDA(cudaMemcpyAsync(d_lam, h_lam, m * 4, cudaMemcpyHostToDevice, st));
        BCP_CUDA(cudaMemcpyAsync(d_state, h_state, m * 4, cudaMemcpyHostToDevice, st));
        BCP_CUDA(cudaMemcpyAsync(d_idx, h_idx, m * 4 * (size_t)T.P, cudaMemcpyHostToDevice, st));
        if (T.pf.q >= 0) BCP_CUDA(launch_energy_pf(T, (long long)m, lam ? d_lam : nullptr, d_state, d_idx, d_out, 0, st));
        else neglogp_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(T, (long long)m, lam ? d_lam : nullptr, d_state, d_idx, d_out);
        BCP_LAUNCHED(1);
        BCP_CUDA(cudaGetLastError());
        BCP_CUDA(cudaMemcpyAsync(h_out, d_out, m * 8, cudaMemcpyDeviceToHost, st));
        BCP_CUDA(cudaStreamSynchronize(st));
        memcpy(energy_out + i0, h_out, m * 8);
    }
    return BCP_OK;
}
