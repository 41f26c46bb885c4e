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
__device__ __forceinline__ double sel3(int d, double m, double z, double p) {   // d in {-1,0,1}
    return d < 0 ? m : (d > 0 ? p : z);
}

// state-move words of the NEXT step, filled one step ahead (ping-pong: no register copies)
template <int R>
struct Prefetch {
    double C;
    double2 rec[R];
    double W[5];
};

// a / b with y = RN(1/b): bit-identical to IEEE division for the magnitudes SamplerTables::div_safe admits
// (sampler_spec.cu:div_markstein); 5 dependent FMA-class operations instead of the ~20-instruction sequence
__device__ __forceinline__ double fast_div(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, y, q);
}

template <int R, bool HG, bool MK>
__global__ void __launch_bounds__(128)
mcmc_fast_kernel(const __grid_constant__ SamplerTables T, const __grid_constant__ ChainBuffers C,
                 const __grid_constant__ LaunchArgs A) {
    constexpr int QG = R - 1;               // index of the gamma restraint (if HG)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_sig = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int4 s_meta[BCP_MAX_RESTRAINTS];     // {n_sigma, sig_off, hist_sigma, 0} by restraint
    if (threadIdx.x < R)
        s_meta[threadIdx.x] = make_int4(T.r[threadIdx.x].n_sigma, T.r[threadIdx.x].sig_off, T.r[threadIdx.x].hist_sigma, 0);
    stage_sigma_table(s_sig, T.sig_tab, T.sig_bytes, &mbar);        // contains a __syncthreads
    // correctly rounded reciprocals of the staged words (used for the 2 sigma^2 entries only)
    double* s_rcp = s_sig + T.sig_bytes / 8;
    if (MK) {
        for (int k = threadIdx.x; k < T.sig_bytes / 8; k += blockDim.x) s_rcp[k] = __drcp_rn(s_sig[k]);
        __syncthreads();
    }

    const long long n = C.n_chains;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;

    const int lam = C.lam[c];
    const int grp = C.group[c];
    const unsigned long long chain_id = C.chain_id0 + (unsigned long long)c;
    const double* __restrict__ energy = T.energy + (size_t)lam * T.S;
    const double logZ = T.logZ[lam];
    const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(T.rec);
    const double* __restrict__ gsse = HG ? T.gsse_halo : nullptr;     // rows of G + 4 (periodic halo)
    const int G = HG ? T.r[QG].n_gamma : 1;           // >= 2 (checked on the host)
    const int GH = G + 4;
    unsigned long long* g_state_counts = C.state_counts + (size_t)grp * T.S;
    unsigned long long* g_param_counts = C.param_counts + (size_t)grp * T.HB;
    const int S = T.S;
    const uint32_t thresh = T.nuis_thresh;
    const int s_end = (int)(A.s_end - A.s_begin);                    // launch-local step counter
    const long long burn_l = A.burn - A.s_begin;                     // may be negative
    const int burn = burn_l < 0 ? 0 : (burn_l > s_end ? s_end : (int)burn_l);
    const unsigned long long ctr0 = A.step0 + (unsigned long long)A.s_begin;
    const uint32_t seed_lo = (uint32_t)C.seed, seed_hi = (uint32_t)(C.seed >> 32);
    const uint32_t ch_lo = (uint32_t)chain_id, ch_hi = (uint32_t)(chain_id >> 32);

    // the 20 Philox round keys are launch constants: kept in registers (opaque to the compiler, which would
    // otherwise re-derive them from the seed in every round)
    uint32_t rk0[10], rk1[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        rk0[i] = seed_lo + (uint32_t)i * BCP_PHILOX_W0;
        rk1[i] = seed_hi + (uint32_t)i * BCP_PHILOX_W1;
        asm volatile("" : "+r"(rk0[i]), "+r"(rk1[i]));
    }
    auto philox = [&](unsigned long long ctr) {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = ch_lo, c3 = ch_hi;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint64_t p0 = (uint64_t)BCP_PHILOX_M0 * c0;
            const uint64_t p1 = (uint64_t)BCP_PHILOX_M1 * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk0[i];
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk1[i];
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        }
        Philox4 r;
        r.w0 = c0; r.w1 = c1; r.w2 = c2; r.w3 = c3;
        return r;
    };

    // ---- chain state -------------------------------------------------------------------
    int state = C.state[c];
    double E = C.E[c];
    int isg[R];
    int igm = 0;
    double csse[R], cref[R];            // csse[QG] unused when HG (3-wide window instead)
    double W0 = 0.0, W1 = 0.0, W2 = 0.0;
    int since_sg[R];
    int since_gm = 0, since_state = 0;
    unsigned int acc_r[R];
    unsigned int acc_state = 0, acc_total = 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        isg[q] = C.isg[(size_t)q * n + c];
        const double2 v = rec2[(size_t)state * R + q];
        csse[q] = v.x;
        cref[q] = v.y;
        since_sg[q] = 0;
        acc_r[q] = 0;
    }
    if (HG) {
        igm = C.igm[(size_t)QG * n + c];
        const double* row = gsse + (size_t)state * GH + igm + 1;      // halo: [igm-1, igm, igm+1]
        W0 = row[0];
        W1 = row[1];
        W2 = row[2];
    }
    double Cst = __dadd_rn(energy[state], logZ);
    int pshift = 0;

    int next_snap = -1;
    long long snap_idx = 0;
    if (A.traj_every > 0) {
        const long long m = A.s_begin <= A.burn ? 0 : (A.s_begin - A.burn + A.traj_every - 1) / A.traj_every;
        snap_idx = m;
        const long long ns = A.burn + m * A.traj_every - A.s_begin;
        next_snap = ns < s_end ? (int)ns : -1;
    }

    // issue the loads a state move to state i2 will need (igm = gamma index before the
    // current step resolves; the 5-wide window covers a +-1 change)
    auto prefetch = [&](Prefetch<R>& p, int i2) {
        p.C = __ldg(energy + i2);
#pragma unroll
        for (int q = 0; q < R; ++q) p.rec[q] = __ldg(rec2 + (size_t)i2 * R + q);
        if (HG) {
            const double* row = gsse + (size_t)i2 * GH + igm;            // halo: [igm-2 .. igm+2], no wrap
#pragma unroll
            for (int k = 0; k < 5; ++k) p.W[k] = __ldg(row + k);
        }
    };

    // one Metropolis step: consumes `use` (if this step is a state move), fills `fill` for the
    // next step, reads the RNG block w and produces the next one in wn
    auto step = [&](const int s, const Philox4& w, Philox4& wn, const Prefetch<R>& use, Prefetch<R>& fill) {
        // ---- decode the proposal ---------------------------------------------------------
        const bool nuis = w.w0 < thresh;
        const uint64_t pr = (uint64_t)w.w1 * (uint32_t)(nuis ? R : S);
        const int hi = (int)(pr >> 32);
        const int pick = nuis ? hi : -1;
        const int nstate = nuis ? state : hi;
        const uint64_t z1 = (uint64_t)(uint32_t)pr * 3u;
        const int dsg = (int)(z1 >> 32) - 1;
        const uint64_t z2 = (uint64_t)(uint32_t)z1 * 3u;
        const bool pg = HG && pick == QG;
        const int dgm = pg ? (int)(z2 >> 32) - 1 : 0;
        const double u2 = u2_from_words(w.w2, w.w3);

        const int4 meta = s_meta[nuis ? hi : 0];       // {n_sigma, sig_off, hist_sigma}
        int sg_old = 0, since_sel = 0;
#pragma unroll
        for (int q = 0; q < R; ++q)
            if (q == pick) { sg_old = isg[q]; since_sel = since_sg[q]; }
        const int sgn = wrap1(sg_old + dsg, meta.x);   // meaningful for nuisance lanes only
        const int gmn = HG ? wrap1(igm + dgm, G) : 0;

        // ---- operands of the R terms -----------------------------------------------------
        double num[R], ref[R], av[R], dv[R], yv[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int sq = (q == pick) ? sgn : isg[q];
            av[q] = s_sig[T.r[q].sig_off + sq];
            dv[q] = s_sig[T.r[q].sig_off + T.r[q].n_sigma + sq];
            yv[q] = MK ? s_rcp[T.r[q].sig_off + T.r[q].n_sigma + sq] : 0.0;
            num[q] = nuis ? csse[q] : use.rec[q].x;
            ref[q] = nuis ? cref[q] : use.rec[q].y;
        }
        double nW0 = 0.0, nW2 = 0.0;
        if (HG) {
            num[QG] = nuis ? sel3(dgm, W0, W1, W2) : sel3(pshift, use.W[1], use.W[2], use.W[3]);
            nW0 = sel3(pshift, use.W[0], use.W[1], use.W[2]);
            nW2 = sel3(pshift, use.W[2], use.W[3], use.W[4]);
        }
        const double Cn = nuis ? Cst : __dadd_rn(use.C, logZ);

        // ---- next step: RNG and, if it is a state move, its table words -------------------
        {
            const unsigned long long ctr = ctr0 + (unsigned long long)(s + 1);
            wn = philox(ctr);
            if (!(wn.w0 < thresh)) prefetch(fill, (int)(((uint64_t)wn.w1 * (uint32_t)S) >> 32));
        }

        // ---- energy of the proposal, reference fp64 order ---------------------------------
        double En = Cn;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            double t = __dadd_rn(av[q], MK ? fast_div(num[q], dv[q], yv[q]) : __ddiv_rn(num[q], dv[q]));
            t = __dadd_rn(t, T.r[q].cnorm);
            t = __dsub_rn(t, ref[q]);
            En = __dadd_rn(En, t);
        }
        const bool acc = (En < E) || metropolis_accept(__dsub_rn(E, En), u2);

        // ---- update (predicated; only the histogram flushes branch) -----------------------
        const bool acc_n = acc && nuis, acc_s = acc && !nuis;
        E = acc ? En : E;
        acc_total += acc ? 1u : 0u;
        acc_state += acc_s ? 1u : 0u;
        const bool sg_changed = acc_n && sgn != sg_old;
        if (sg_changed) {
            const int lo = since_sel > burn ? since_sel : burn;
            if (s > lo) atomicAdd(g_param_counts + meta.z + sg_old, (unsigned long long)(s - lo));
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const bool mine = q == pick;
            acc_r[q] += (acc_n && mine) ? 1u : 0u;
            isg[q] = (sg_changed && mine) ? sgn : isg[q];
            since_sg[q] = (sg_changed && mine) ? s : since_sg[q];
            csse[q] = acc_s ? use.rec[q].x : csse[q];
            cref[q] = acc_s ? use.rec[q].y : cref[q];
        }
        Cst = acc_s ? Cn : Cst;
        pshift = 0;
        if (HG) {
            const bool gm_changed = acc_n && dgm != 0;
            if (gm_changed) {
                const int lo = since_gm > burn ? since_gm : burn;
                if (s > lo) atomicAdd(g_param_counts + T.r[QG].hist_gamma + igm, (unsigned long long)(s - lo));
                // slide the window and fetch the new edge
                const double edge = __ldg(gsse + (size_t)state * GH + (gmn + 2 + dgm));
                if (dgm > 0) { W0 = W1; W1 = W2; W2 = edge; }
                else { W2 = W1; W1 = W0; W0 = edge; }
                since_gm = s;
                igm = gmn;
                pshift = dgm;
            }
            W0 = acc_s ? nW0 : W0;
            W1 = acc_s ? num[QG] : W1;
            W2 = acc_s ? nW2 : W2;
        }
        if (acc_s && nstate != state) {
            const int lo = since_state > burn ? since_state : burn;
            if (s > lo) atomicAdd(g_state_counts + state, (unsigned long long)(s - lo));
            since_state = s;
            state = nstate;
        }

        if (s >= burn) {
            if (A.state_trace && c < A.trace_n) A.state_trace[(size_t)(A.s_begin + s - A.burn) * A.trace_n + c] = state;
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
    };

    // ---- software pipeline: prologue, then two steps per iteration (ping-pong buffers) ------
    Philox4 wa = philox(ctr0), wb;
    Prefetch<R> pa, pb;
    pa.C = 0.0; pb.C = 0.0;
#pragma unroll
    for (int q = 0; q < R; ++q) { pa.rec[q] = make_double2(0.0, 0.0); pb.rec[q] = make_double2(0.0, 0.0); }
#pragma unroll
    for (int k = 0; k < 5; ++k) { pa.W[k] = 0.0; pb.W[k] = 0.0; }
    if (!(wa.w0 < thresh)) prefetch(pa, (int)(((uint64_t)wa.w1 * (uint32_t)S) >> 32));
    int s = 0;
    for (; s + 1 < s_end; s += 2) {
        step(s, wa, wb, pa, pb);
        step(s + 1, wb, wa, pb, pa);
    }
    if (s < s_end) step(s, wa, wb, pa, pb);

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

template <int R, bool HG, bool MK>
static cudaError_t launch_one_mk(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid, int block,
                                 cudaStream_t st) {
    const int smem = T.sig_bytes * (MK ? 2 : 1);
    cudaError_t e = cudaFuncSetAttribute(mcmc_fast_kernel<R, HG, MK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    mcmc_fast_kernel<R, HG, MK><<<grid, block, smem, st>>>(T, C, A);
    return cudaGetLastError();
}
template <int R, bool HG>
static cudaError_t launch_one(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid, int block,
                              cudaStream_t st) {
    // the reciprocal-table division needs every operand in the safe range checked by bcp_create
    return T.div_safe ? launch_one_mk<R, HG, true>(T, C, A, grid, block, st) : launch_one_mk<R, HG, false>(T, C, A, grid, block, st);
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
