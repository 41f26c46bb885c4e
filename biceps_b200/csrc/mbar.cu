// K6 + K7: snapshot rescoring under every lambda (u_kn) and MBAR reweighting to the BICePs
// score and populations.
//
// Replaces
//   Analysis.MBAR_analysis u_kln build      /root/reference/biceps/Analysis.py:145-185
//   pymbar.MBAR(u_kln, N_k)                  called at /root/reference/biceps/Analysis.py:192
//   compute_free_energy_differences          :201   (uncertainty_method='approximate': Theta = W^T W)
//   compute_expectations(indicator)          :217-223 (populations P_i(l) and their uncertainties)
// pymbar (>= 4.0.1, requirements.txt:27) is a third-party dependency absent from the reference
// tree; the MBAR equations of Shirts & Chodera (JCP 129, 124105, 2008) are restated here:
//   logden_n = ln sum_k N_k exp(f_k - u_kn)          W_nk = exp(f_k - u_kn - logden_n)
//   self-consistent:  f_k <- f_k - ln sum_n W_nk      Newton on the convex objective
//   g_k = N_k (sum_n W_nk - 1),  H_kl = delta_kl N_k s_k - N_k N_l sum_n W_nk W_nl
//
// One streaming pass per iteration: u_kn is read once (8*K*N bytes, coalesced over n), the
// per-sample log-sum-exp over k is done in registers/shared memory, and s_k / W^T W are
// reduced deterministically (per-CTA partials, fixed-order finish).
#include "common.cuh"
#include "tables.cuh"
#include "sampler_common.cuh"

#include <math.h>
#include <algorithm>
#include <vector>

namespace bcp {

constexpr int MB_TILE = 256;             // samples per tile == threads per CTA
constexpr int MB_LD = MB_TILE + 1;       // padded leading dimension (bank-conflict-free pair sums)
constexpr int MB_MAXK = 96;

// exp(x) for the shifted exponents of a log-sum-exp (x <= 0, -inf allowed): argument reduction x = k ln2 + r,
// degree-13 Taylor polynomial on |r| <= ln2 / 2 (truncation 4e-18), scaling by 2^k through the exponent field.
// Branch-free (libdevice's exp carries a slow-path branch per call, which splits the unrolled K exponentials of a
// sample into K basic blocks and keeps them from overlapping) and 22 instead of ~33 instructions; <= 1 ulp away
// from libdevice's result.  Arguments below -708 (results below 3e-308, against a sum >= 1) are clamped.
__device__ __forceinline__ double exp_stream(double x) {
    x = fmax(x, -708.0);
    const double t = fma(x, 0x1.71547652b82fep+0, 6755399441055744.0);
    const double k = t - 6755399441055744.0;
    double r = fma(k, -0x1.62e42fefa39efp-1, x);
    r = fma(k, -0x1.abc9e3b39803fp-56, r);
    double p = 0x1.6124613a86d09p-33;
    p = fma(p, r, 0x1.1eed8eff8d898p-29);
    p = fma(p, r, 0x1.ae64567f544e4p-26);
    p = fma(p, r, 0x1.27e4fb7789f5cp-22);
    p = fma(p, r, 0x1.71de3a556c734p-19);
    p = fma(p, r, 0x1.a01a01a01a01ap-16);
    p = fma(p, r, 0x1.a01a01a01a01ap-13);
    p = fma(p, r, 0x1.6c16c16c16c17p-10);
    p = fma(p, r, 0x1.1111111111111p-7);
    p = fma(p, r, 0x1.5555555555555p-5);
    p = fma(p, r, 0x1.5555555555555p-3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ki = __double2loint(t);                      // low word of k + 1.5 * 2^52 == k (two's complement)
    return p * __hiloint2double((ki + 1023) << 20, 0);
}

// The same with a 64-entry table: x = (64 e + j) ln2 / 64 + r, |r| <= ln2 / 128, exp(x) = 2^e * 2^(j/64) * p(r) with a
// degree-5 Taylor polynomial (truncation r^6 / 720 <= 3.5e-17).  11 fp64 instructions + one shared-memory load instead
// of 22: the streaming passes are bound by the fp64 pipe (K exponentials per sample), not by HBM, with exp_stream
// (ncu: fp64 pipe 67 % active at 43 % of the HBM peak).  tab[j] = 2^(j/64) in shared memory (exp_tab_init); ~2 ulp.
__device__ __forceinline__ void exp_tab_init(double* tab) {
    if (threadIdx.x < 64) tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 64.0));
    __syncthreads();
}
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
    x = fmax(x, -708.0);
    const double t = fma(x, 0x1.71547652b82fep+6, 6755399441055744.0);      // 64 / ln2
    const double k = t - 6755399441055744.0;
    double r = fma(k, -0x1.62e42fee00000p-7, x);                            // ln2 / 64, high part (k * hi is exact)
    r = fma(k, -0x1.a39ef35793c76p-39, r);
    double p = fma(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);          // 1/120, 1/24
    p = fma(p, r, 0x1.5555555555555p-3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ki = __double2loint(t);                      // low word of k + 1.5 * 2^52 == k (two's complement)
    const double s = tab[ki & 63];
    return p * __hiloint2double(__double2hiint(s) + ((ki >> 6) << 20), __double2loint(s));
}

// a[k] = f_k + ln N_k (or -inf if N_k == 0);  rn[k] = 1/N_k
__device__ __forceinline__ void sample_weights(const double* __restrict__ u, long long N, long long n, int K,
                                               const double* __restrict__ a, double* __restrict__ Ws, int t) {
    // pass 1: shifted exponents and their max
    double mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
        const double v = a[k] - __ldcs(u + (size_t)k * N + n);
        Ws[k * MB_LD + t] = v;
        mx = fmax(mx, v);
    }
    double sum = 0.0;
    for (int k = 0; k < K; ++k) {
        const double e = exp_stream(Ws[k * MB_LD + t] - mx);
        Ws[k * MB_LD + t] = e;
        sum += e;
    }
    const double inv = 1.0 / sum;
    // W_nk = exp(f_k - u_kn - logden) = e_k / (N_k * sum)  ->  store N_k W_nk = e_k/sum first
    for (int k = 0; k < K; ++k) Ws[k * MB_LD + t] *= inv;
}

// partial[block][p]: p < K           -> sum_n W_nk
//                    p >= K (pairs)  -> sum_n W_nk W_nl for k <= l, row-major upper triangle
template <bool HESS>
__global__ void __launch_bounds__(MB_TILE)
mbar_pass_kernel(const double* __restrict__ u, long long N, int K, const double* __restrict__ a_in,
                 const double* __restrict__ rn_in, double* __restrict__ partial, int n_acc) {
    extern __shared__ double smem[];
    double* Ws = smem;                       // [K][MB_LD]
    double* a = Ws + (size_t)K * MB_LD;      // [K]
    double* rn = a + K;                      // [K]
    short* pk = reinterpret_cast<short*>(rn + K);   // pair index tables [n_acc] x 2
    short* pl = pk + n_acc;
    const int t = threadIdx.x;
    for (int k = t; k < K; k += MB_TILE) { a[k] = a_in[k]; rn[k] = rn_in[k]; }
    for (int p = t; p < n_acc; p += MB_TILE) {
        if (p < K) { pk[p] = (short)p; pl[p] = -1; }
        else {
            int q = p - K, k = 0;
            while (q >= K - k) { q -= K - k; ++k; }
            pk[p] = (short)k; pl[p] = (short)(k + q);
        }
    }
    __syncthreads();
    constexpr int MAX_ACC = (MB_MAXK + MB_MAXK * (MB_MAXK + 1) / 2 + MB_TILE - 1) / MB_TILE;
    double acc[MAX_ACC];
#pragma unroll
    for (int i = 0; i < MAX_ACC; ++i) acc[i] = 0.0;

    const long long n_tiles = (N + MB_TILE - 1) / MB_TILE;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile * MB_TILE + t;
        if (n < N) {
            sample_weights(u, N, n, K, a, Ws, t);
            for (int k = 0; k < K; ++k) Ws[k * MB_LD + t] *= rn[k];
        } else {
            for (int k = 0; k < K; ++k) Ws[k * MB_LD + t] = 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MAX_ACC; ++i) {
            const int p = t + i * MB_TILE;
            if (p < n_acc) {
                const double* rk = Ws + pk[p] * MB_LD;
                double s0 = 0.0, s1 = 0.0;
                if (pl[p] < 0) {
                    for (int j = 0; j < MB_TILE; j += 2) { s0 += rk[j]; s1 += rk[j + 1]; }
                } else if (HESS) {
                    const double* rl = Ws + pl[p] * MB_LD;
                    for (int j = 0; j < MB_TILE; j += 2) { s0 += rk[j] * rl[j]; s1 += rk[j + 1] * rl[j + 1]; }
                }
                acc[i] += s0 + s1;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MAX_ACC; ++i) {
        const int p = t + i * MB_TILE;
        if (p < n_acc) partial[(size_t)blockIdx.x * n_acc + p] = acc[i];
    }
}

// Streaming pass for small K (the BICePs case: a handful of lambda values): the K reduced potentials of a
// sample, its log-sum-exp and weights stay in registers, and every thread accumulates its own partial s_k
// and (HESS) upper triangle of W^T W over its samples -- no barrier inside the stream.  The accumulators
// (77 doubles at K = 11) leave room for only 8 warps per SM, far too few plain loads in flight for HBM, so
// u_kn is staged through shared memory with cp.async: every thread copies the K x 16 bytes of ITS two
// samples for MB_STAGES tiles ahead (data in flight holds no registers: 3 x 256 x K x 16 B = 135 KB per SM
// at K = 11) and later reads back what it copied itself, so no block-level synchronisation is needed.
// u_kn is read exactly once.  The partials are reduced once at the end: xor butterfly inside the warp,
// fixed-order sum over the warps.  Same arithmetic per sample as sample_weights().
constexpr int MB_REGK = 12;

// MB_STAGES tiles in flight per CTA: 3 for the Hessian pass (255 registers: one CTA per SM), 2 for the plain
// pass (128 registers, 90 KB at K = 11: two CTAs per SM, twice the warps to overlap the exponentials)
template <int K, bool HESS, int MB_STAGES>
__global__ void __launch_bounds__(256)
mbar_pass_reg_kernel(const double* __restrict__ u, long long N, const double* __restrict__ a_in,
                     const double* __restrict__ rn_in, double* __restrict__ partial) {
    constexpr int NP = HESS ? K * (K + 1) / 2 : 0;
    extern __shared__ __align__(16) double stage_buf[];          // [MB_STAGES][K][256] x double2
    __shared__ double red[8][K + (HESS ? K * (K + 1) / 2 : 0)];
    __shared__ double exp2_tab[64];
    exp_tab_init(exp2_tab);
    double a[K], rn[K], accs[K], accm[NP > 0 ? NP : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) { a[k] = a_in[k]; rn[k] = rn_in[k]; accs[k] = 0.0; }
#pragma unroll
    for (int p = 0; p < NP; ++p) accm[p] = 0.0;

    auto one_sample = [&](const double* x) {                     // x[k] = u_kn of one sample
        double w[K];
        double mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            w[k] = a[k] - x[k];
            mx = fmax(mx, w[k]);
        }
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            w[k] = exp_tab(w[k] - mx, exp2_tab);
            sum += w[k];
        }
        const double inv = 1.0 / sum;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            w[k] = (w[k] * inv) * rn[k];
            accs[k] += w[k];
        }
        if (HESS) {
            int p = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int l = k; l < K; ++l) accm[p++] += w[k] * w[l];
        }
    };

    // tiles of 512 samples (2 per thread); rows of u_kn are 16-byte aligned for every k only if N is even
    const long long n_tiles = N / 512;
    const bool aligned = (N & 1) == 0;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage_buf) + 16u * threadIdx.x;
    auto issue = [&](long long tile, int st) {
        if (tile < n_tiles) {
            const double* src = u + tile * 512 + 2 * threadIdx.x;
            const uint32_t dst = sbase + (uint32_t)st * (K * 256u * 16u);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (aligned) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 4096u), "l"(src + (size_t)k * N) : "memory");
                } else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + k * 4096u), "l"(src + (size_t)k * N) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + k * 4096u + 8u), "l"(src + (size_t)k * N + 1) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long long tile = blockIdx.x;
#pragma unroll
    for (int st = 0; st < MB_STAGES - 1; ++st) issue(tile + (long long)st * gridDim.x, st);
    int st = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        issue(tile + (long long)(MB_STAGES - 1) * gridDim.x, (st + MB_STAGES - 1) % MB_STAGES);
        asm volatile("cp.async.wait_group %0;" ::"n"(MB_STAGES - 1) : "memory");
        const double2* buf = reinterpret_cast<const double2*>(stage_buf) + (size_t)st * (K * 256) + threadIdx.x;
        double x0[K], x1[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double2 v = buf[k * 256];
            x0[k] = v.x; x1[k] = v.y;
        }
        one_sample(x0);
        one_sample(x1);
        st = (st + 1) % MB_STAGES;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // the last N % 512 samples: plain streaming loads
    for (long long n = n_tiles * 512 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
        double x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = __ldcs(u + (size_t)k * N + n);
        one_sample(x);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = accs[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][k] = v;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        double v = accm[p];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][K + p] = v;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < K + NP; p += blockDim.x) {
        double v = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) v += red[wv][p];
        partial[(size_t)blockIdx.x * (K + NP) + p] = v;
    }
}

// returns the number of CTAs launched (= rows of `partial`)
template <int K>
static int launch_pass_reg(bool hess, int grid, int grid_plain, cudaStream_t st, const double* u, long long N, const double* a, const double* rn,
                           double* partial) {
    if (hess) {
        const int smem = 3 * K * 256 * 16;
        cudaFuncSetAttribute(mbar_pass_reg_kernel<K, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        mbar_pass_reg_kernel<K, true, 3><<<grid, 256, smem, st>>>(u, N, a, rn, partial);
        return grid;
    }
    const int smem = 2 * K * 256 * 16;
    cudaFuncSetAttribute(mbar_pass_reg_kernel<K, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mbar_pass_reg_kernel<K, false, 2><<<grid_plain, 256, smem, st>>>(u, N, a, rn, partial);
    return grid_plain;
}

// fixed-order reduction of the per-CTA partials: one warp per accumulator (lane-strided partial sums, then
// the xor butterfly) -- deterministic for a given grid
__global__ void mbar_finish_kernel(const double* __restrict__ partial, int n_blocks, int n_acc, double* __restrict__ out) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_acc) return;
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += partial[(size_t)b * n_acc + p];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[p] = s;
}

// populations: P[l][s] += W_nl, Q[l][s] += W_nl^2 for the state s of sample n (fp64 atomics).
// Consecutive samples are consecutive snapshots of one chain and mostly share their state, and a peaked
// posterior sends almost every sample to a handful of states: a warp whose 32 samples agree reduces its
// weights with shuffles first and issues one atomic per lambda instead of 32 colliding ones.
__global__ void __launch_bounds__(MB_TILE)
mbar_pop_kernel(const double* __restrict__ u, long long N, int K, const double* __restrict__ a_in,
                const double* __restrict__ rn_in, const int* __restrict__ state, int S, double* __restrict__ P,
                double* __restrict__ Q) {
    extern __shared__ double smem[];
    double* Ws = smem;
    double* a = Ws + (size_t)K * MB_LD;
    double* rn = a + K;
    const int t = threadIdx.x;
    for (int k = t; k < K; k += MB_TILE) { a[k] = a_in[k]; rn[k] = rn_in[k]; }
    __syncthreads();
    const long long n_tiles = (N + MB_TILE - 1) / MB_TILE;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile * MB_TILE + t;
        const bool valid = n < N;
        if (valid) sample_weights(u, N, n, K, a, Ws, t);
        const int s = valid ? state[n] : -1;
        const int s0 = __shfl_sync(0xffffffffu, s, 0);
        const bool uniform = __all_sync(0xffffffffu, s == s0 || !valid);
        if (s0 < 0) continue;                               // the whole warp is past the end (warp-uniform)
        if (uniform) {
            for (int k = 0; k < K; ++k) {
                double w = valid ? Ws[k * MB_LD + t] * rn[k] : 0.0;
                double w2 = w * w;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    w += __shfl_xor_sync(0xffffffffu, w, off);
                    w2 += __shfl_xor_sync(0xffffffffu, w2, off);
                }
                if ((t & 31) == 0) {
                    atomicAdd(P + (size_t)k * S + s0, w);
                    atomicAdd(Q + (size_t)k * S + s0, w2);
                }
            }
        } else if (valid) {
            for (int k = 0; k < K; ++k) {
                const double w = Ws[k * MB_LD + t] * rn[k];
                atomicAdd(P + (size_t)k * S + s, w);
                atomicAdd(Q + (size_t)k * S + s, w * w);
            }
        }
    }
}

// populations for small K: the weights of a sample stay in registers (same arithmetic as sample_weights), a
// warp whose 32 samples share their state reduces with shuffles and issues one atomic pair per lambda
template <int K>
__global__ void __launch_bounds__(256)
mbar_pop_reg_kernel(const double* __restrict__ u, long long N, const double* __restrict__ a_in,
                    const double* __restrict__ rn_in, const int* __restrict__ state, int S, double* __restrict__ P,
                    double* __restrict__ Q) {
    __shared__ double exp2_tab[64];
    exp_tab_init(exp2_tab);
    double a[K], rn[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { a[k] = a_in[k]; rn[k] = rn_in[k]; }
    const long long n_warp_tiles = (N + 31) / 32;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long wt = warp0; wt < n_warp_tiles; wt += n_warps) {
        const long long n = wt * 32 + lane;
        const bool valid = n < N;
        double w[K];
        double mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            w[k] = valid ? a[k] - __ldcs(u + (size_t)k * N + n) : 0.0;
            mx = fmax(mx, w[k]);
        }
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            w[k] = exp_tab(w[k] - mx, exp2_tab);
            sum += w[k];
        }
        const double inv = 1.0 / sum;
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = valid ? (w[k] * inv) * rn[k] : 0.0;
        const int s = valid ? state[n] : -1;
        const int s0 = __shfl_sync(0xffffffffu, s, 0);
        const bool uniform = __all_sync(0xffffffffu, s == s0 || !valid);
        if (uniform) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double x = w[k], x2 = w[k] * w[k];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    x += __shfl_xor_sync(0xffffffffu, x, off);
                    x2 += __shfl_xor_sync(0xffffffffu, x2, off);
                }
                if (lane == 0) {
                    atomicAdd(P + (size_t)k * S + s0, x);
                    atomicAdd(Q + (size_t)k * S + s0, x2);
                }
            }
        } else if (valid) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                atomicAdd(P + (size_t)k * S + s, w[k]);
                atomicAdd(Q + (size_t)k * S + s, w[k] * w[k]);
            }
        }
    }
}

// K6: u[l][n] = ((energy[l][state_n] + logZ[l]) + t_0) + t_1 + ...  (PosteriorSampler.neglogP order);
// the restraint terms do not depend on lambda and are evaluated once per snapshot.
__global__ void __launch_bounds__(256)
ukn_kernel(const SamplerTables T, const double* __restrict__ energy_t, long long N, const int* __restrict__ state,
           const int* __restrict__ indices, double* __restrict__ u) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int st = state[n];
    double t[BCP_MAX_RESTRAINTS];
    int p = 0;
#pragma unroll
    for (int q = 0; q < BCP_MAX_RESTRAINTS; ++q) {
        if (q < T.R) {
            const RestraintDev& rd = T.r[q];
            const int sg = indices[(size_t)n * T.P + p++];
            const int gm = rd.has_gamma ? indices[(size_t)n * T.P + p++] : 0;
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
        u[(size_t)l * N + n] = E;
    }
}

// ---------------------------------------------------------------------------------------
static size_t pass_smem(int K, int n_acc) { return ((size_t)K * MB_LD + 2 * K) * sizeof(double) + 2 * (size_t)n_acc * sizeof(short) + 16; }

struct MbarWork {
    int device = 0, K = 0, sms = 148;
    long long N = 0;
    int n_acc_h = 0, grid = 0, grid_plain = 0;
    double *d_a = nullptr, *d_rn = nullptr, *d_partial = nullptr, *d_out = nullptr;
    std::vector<double> lnN, rn;
    cudaStream_t st = nullptr;      // workspace is stream-ordered (cudaMallocAsync pool)
    ~MbarWork() { cudaFreeAsync(d_a, st); cudaFreeAsync(d_rn, st); cudaFreeAsync(d_partial, st); cudaFreeAsync(d_out, st); }
};

static int mbar_work_init(MbarWork& w, int device, int K, long long N, const int64_t* N_k, cudaStream_t st) {
    w.device = device; w.K = K; w.N = N; w.sms = num_sms(device); w.st = st;
    w.n_acc_h = K + K * (K + 1) / 2;
    const long long tiles = (N + MB_TILE - 1) / MB_TILE;
    w.grid = (int)std::min<long long>(tiles, (long long)w.sms * 4);
    if (K <= MB_REGK) {      // persistent CTAs of the staged small-K kernel: as many per SM as their shared memory allows
        const int smem = 3 * K * 256 * 16;                  // Hessian pass: 3 stages
        const int per_sm = std::max(1, std::min(4, 220 * 1024 / smem));
        w.grid = (int)std::min<long long>((N + 511) / 512, (long long)w.sms * per_sm);
        const int smem2 = 2 * K * 256 * 16;                 // plain pass: 2 stages, 128 registers -> 2 CTAs per SM
        const int per_sm2 = std::max(1, std::min(2, 220 * 1024 / smem2));
        w.grid_plain = (int)std::min<long long>((N + 511) / 512, (long long)w.sms * per_sm2);
    }
    if (w.grid < 1) w.grid = 1;
    if (w.grid_plain < 1) w.grid_plain = w.grid;
    w.lnN.resize(K); w.rn.resize(K);
    for (int k = 0; k < K; ++k) {
        w.lnN[k] = N_k[k] > 0 ? log((double)N_k[k]) : -INFINITY;
        w.rn[k] = N_k[k] > 0 ? 1.0 / (double)N_k[k] : 0.0;
    }
    BCP_CUDA(bcp::pool_malloc((void**)&w.d_a, K * 8, st)); BCP_CUDA(bcp::pool_malloc((void**)&w.d_rn, K * 8, st));
    BCP_CUDA(bcp::pool_malloc((void**)&w.d_partial, std::max((size_t)w.grid * w.n_acc_h, (size_t)w.grid_plain * K) * 8, st));
    BCP_CUDA(bcp::pool_malloc((void**)&w.d_out, (size_t)w.n_acc_h * 8, st));
    BCP_CUDA(cudaMemcpyAsync(w.d_rn, w.rn.data(), K * 8, cudaMemcpyHostToDevice, st));
    const size_t sm = pass_smem(K, w.n_acc_h);
    BCP_CUDA(cudaFuncSetAttribute(mbar_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    BCP_CUDA(cudaFuncSetAttribute(mbar_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    BCP_CUDA(cudaFuncSetAttribute(mbar_pop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    return BCP_OK;
}

// one streaming pass at free energies f: s[K] = sum_n W_nk and (hess) M[K][K] = W^T W
static int mbar_pass(MbarWork& w, const double* d_u, const double* f, bool hess, double* s, double* M, cudaStream_t st) {
    const int K = w.K;
    std::vector<double> a(K);
    for (int k = 0; k < K; ++k) a[k] = f[k] + w.lnN[k];
    BCP_CUDA(cudaMemcpyAsync(w.d_a, a.data(), K * 8, cudaMemcpyHostToDevice, st));
    const int n_acc = hess ? w.n_acc_h : K;
    const size_t sm = pass_smem(K, n_acc);
    int n_blocks = w.grid;
    if (K <= MB_REGK) {
        switch (K) {
#define BCP_PASS(k) case k: n_blocks = launch_pass_reg<k>(hess, w.grid, w.grid_plain, st, d_u, w.N, w.d_a, w.d_rn, w.d_partial); break;
            BCP_PASS(1) BCP_PASS(2) BCP_PASS(3) BCP_PASS(4) BCP_PASS(5) BCP_PASS(6)
            BCP_PASS(7) BCP_PASS(8) BCP_PASS(9) BCP_PASS(10) BCP_PASS(11) BCP_PASS(12)
#undef BCP_PASS
        }
    }
    else if (hess) mbar_pass_kernel<true><<<w.grid, MB_TILE, sm, st>>>(d_u, w.N, K, w.d_a, w.d_rn, w.d_partial, n_acc);
    else mbar_pass_kernel<false><<<w.grid, MB_TILE, sm, st>>>(d_u, w.N, K, w.d_a, w.d_rn, w.d_partial, n_acc);
    mbar_finish_kernel<<<(n_acc * 32 + 127) / 128, 128, 0, st>>>(w.d_partial, n_blocks, n_acc, w.d_out);
    BCP_LAUNCHED(2);
    BCP_CUDA(cudaGetLastError());
    std::vector<double> out(n_acc);
    BCP_CUDA(cudaMemcpyAsync(out.data(), w.d_out, (size_t)n_acc * 8, cudaMemcpyDeviceToHost, st));
    BCP_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < K; ++k) s[k] = out[k];
    if (hess && M) {
        int p = K;
        for (int k = 0; k < K; ++k)
            for (int l = k; l < K; ++l) { M[k * K + l] = out[p]; M[l * K + k] = out[p]; ++p; }
    }
    return BCP_OK;
}

// dense symmetric positive-definite solve (Cholesky), n <= MB_MAXK; returns false if not PD
static bool chol_solve(std::vector<double>& A, std::vector<double>& b, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = A[i * n + j];
            for (int k = 0; k < j; ++k) v -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = v / d;
        }
    }
    for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= A[i * n + k] * b[k]; b[i] = v / A[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double v = b[i]; for (int k = i + 1; k < n; ++k) v -= A[k * n + i] * b[k]; b[i] = v / A[i * n + i]; }
    return true;
}

static double grad_norm(const double* s, const int64_t* N_k, int K) {
    double g = 0.0;
    for (int k = 0; k < K; ++k) { const double v = (double)N_k[k] * (s[k] - 1.0); g += v * v; }
    return sqrt(g);
}

// d_u: device [K][N]; solves for f (f[0] = 0), returns Theta = W^T W at the solution
static int mbar_solve(MbarWork& w, const double* d_u, const int64_t* N_k, double tol, int max_iter, double* f,
                      double* theta, int* iters, cudaStream_t st) {
    const int K = w.K;
    std::vector<double> s(K), M((size_t)K * K), s2(K), M2((size_t)K * K), fn(K);
    for (int k = 0; k < K; ++k) f[k] = 0.0;
    int rc = mbar_pass(w, d_u, f, true, s.data(), M.data(), st);
    if (rc != BCP_OK) return rc;
    int it = 0;
    bool converged = false;
    for (; it < max_iter; ++it) {
        const double g0 = grad_norm(s.data(), N_k, K);
        bool took_newton = false;
        if (it >= 2 && K > 1) {
            // Newton step on f[1:] (f[0] pinned): H d = -g
            const int n = K - 1;
            std::vector<double> H((size_t)n * n), b(n);
            for (int i = 0; i < n; ++i) {
                const int k = i + 1;
                b[i] = -(double)N_k[k] * (s[k] - 1.0);
                for (int j = 0; j < n; ++j) {
                    const int l = j + 1;
                    H[i * n + j] = (k == l ? (double)N_k[k] * s[k] : 0.0) - (double)N_k[k] * (double)N_k[l] * M[k * K + l];
                }
            }
            if (chol_solve(H, b, n)) {
                fn[0] = 0.0;
                for (int i = 0; i < n; ++i) fn[i + 1] = f[i + 1] + b[i];
                rc = mbar_pass(w, d_u, fn.data(), true, s2.data(), M2.data(), st);
                if (rc != BCP_OK) return rc;
                const double g1 = grad_norm(s2.data(), N_k, K);
                if (g1 < g0 || g1 == 0.0) took_newton = true;
            }
        }
        if (!took_newton) {
            // self-consistent step: f_k <- f_k - ln sum_n W_nk, re-pinned to f_0 = 0
            for (int k = 0; k < K; ++k) fn[k] = N_k[k] > 0 ? f[k] - log(s[k]) : f[k];
            const double f0 = fn[0];
            for (int k = 0; k < K; ++k) fn[k] -= f0;
            rc = mbar_pass(w, d_u, fn.data(), true, s2.data(), M2.data(), st);
            if (rc != BCP_OK) return rc;
        }
        double dmax = 0.0, fmaxv = 1.0;
        for (int k = 0; k < K; ++k) { dmax = fmax(dmax, fabs(fn[k] - f[k])); fmaxv = fmax(fmaxv, fabs(fn[k])); }
        for (int k = 0; k < K; ++k) f[k] = fn[k];
        s.swap(s2); M.swap(M2);
        if (dmax <= tol * fmaxv) { converged = true; ++it; break; }
    }
    if (iters) *iters = it;
    // Theta = W^T W (pymbar uncertainty_method='approximate')
    if (theta) for (int i = 0; i < K * K; ++i) theta[i] = M[i];
    if (!converged) { set_error("bcp_mbar: no convergence in %d iterations", max_iter); return BCP_ERR_NOCONV; }
    return BCP_OK;
}

static int mbar_populations(MbarWork& w, const double* d_u, const double* f, const int* d_state, int S, double* d_P,
                            double* d_Q, cudaStream_t st) {
    const int K = w.K;
    std::vector<double> a(K);
    for (int k = 0; k < K; ++k) a[k] = f[k] + w.lnN[k];
    BCP_CUDA(cudaMemcpyAsync(w.d_a, a.data(), K * 8, cudaMemcpyHostToDevice, st));
    BCP_CUDA(cudaMemsetAsync(d_P, 0, (size_t)K * S * 8, st));
    BCP_CUDA(cudaMemsetAsync(d_Q, 0, (size_t)K * S * 8, st));
    if (K <= MB_REGK) {
        const int pgrid = (int)std::min<long long>((w.N + 255) / 256, (long long)w.sms * 8);
        switch (K) {
#define BCP_POP(k) case k: mbar_pop_reg_kernel<k><<<pgrid, 256, 0, st>>>(d_u, w.N, w.d_a, w.d_rn, d_state, S, d_P, d_Q); break;
            BCP_POP(1) BCP_POP(2) BCP_POP(3) BCP_POP(4) BCP_POP(5) BCP_POP(6)
            BCP_POP(7) BCP_POP(8) BCP_POP(9) BCP_POP(10) BCP_POP(11) BCP_POP(12)
#undef BCP_POP
        }
    } else
        mbar_pop_kernel<<<w.grid, MB_TILE, pass_smem(K, 0), st>>>(d_u, w.N, K, w.d_a, w.d_rn, d_state, S, d_P, d_Q);
    BCP_LAUNCHED(1);
    BCP_CUDA(cudaGetLastError());
    return BCP_OK;
}

}  // namespace bcp

using namespace bcp;

extern "C" int bcp_ukn(bcp_handle* h, int64_t n, const int32_t* state, const int32_t* indices, double* u_kn) {
    BCP_REQUIRE(h && state && indices && u_kn, "bcp_ukn: NULL argument");
    if (n <= 0) return BCP_OK;
    const SamplerTables& T = handle_tables(h);
    int rc;
    DeviceGuard guard(handle_device(h));
    std::vector<int32_t> glen(T.P);
    bcp_param_layout(h, nullptr, glen.data());
    for (int64_t i = 0; i < n; ++i) {
        BCP_REQUIRE(state[i] >= 0 && state[i] < T.S, "bcp_ukn: state[%lld] out of range", (long long)i);
        for (int p = 0; p < T.P; ++p)
            BCP_REQUIRE(indices[i * T.P + p] >= 0 && indices[i * T.P + p] < glen[p], "bcp_ukn: indices[%lld][%d] out of range", (long long)i, p);
    }
    int *d_state = nullptr, *d_idx = nullptr;
    double* d_u = nullptr;
    rc = BCP_OK;
    do {
#define STEP(x) if ((x) != cudaSuccess) { set_error("bcp_ukn: %s: %s", #x, cudaGetErrorString(cudaGetLastError())); rc = BCP_ERR_CUDA; break; }
        STEP(dev_alloc(&d_state, (size_t)n)); STEP(cudaMemcpy(d_state, state, n * 4, cudaMemcpyHostToDevice));
        STEP(dev_alloc(&d_idx, (size_t)n * T.P)); STEP(cudaMemcpy(d_idx, indices, (size_t)n * T.P * 4, cudaMemcpyHostToDevice));
        STEP(dev_alloc(&d_u, (size_t)n * T.K));
        if (T.pf.q >= 0) { STEP(launch_energy_pf(T, n, nullptr, d_state, d_idx, d_u, 1, nullptr)); }
        else ukn_kernel<<<(unsigned)((n + 255) / 256), 256>>>(T, handle_energy_t(h), n, d_state, d_idx, d_u);
        BCP_LAUNCHED(1);
        STEP(cudaGetLastError());
        STEP(cudaMemcpy(u_kn, d_u, (size_t)n * T.K * 8, cudaMemcpyDeviceToHost));
#undef STEP
    } while (0);
    cudaFree(d_state); cudaFree(d_idx); cudaFree(d_u);
    return rc;
}

extern "C" int bcp_ukn_dev(bcp_handle* h, int64_t n, const int32_t* d_state, const int32_t* d_indices, double* d_u_kn,
                           void* stream) {
    BCP_REQUIRE(h && d_state && d_indices && d_u_kn, "bcp_ukn_dev: NULL argument");
    if (n <= 0) return BCP_OK;
    const SamplerTables& T = handle_tables(h);
    DeviceGuard guard(handle_device(h));
    if (T.pf.q >= 0) BCP_CUDA(launch_energy_pf(T, n, nullptr, d_state, d_indices, d_u_kn, 1, (cudaStream_t)stream));
    else ukn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, handle_energy_t(h), n, d_state, d_indices, d_u_kn);
    BCP_LAUNCHED(1);
    BCP_CUDA(cudaGetLastError());
    return BCP_OK;
}

static int mbar_finish_outputs(int K, int S, const double* f, const double* theta, const double* P, const double* Q,
                               double* pops, double* dpops) {
    // P_i(l) and the 'approximate' uncertainty: W_nA = W_nl A_n / P_i(l)
    for (int i = 0; i < S; ++i)
        for (int l = 0; l < K; ++l) {
            const double mu = P[(size_t)l * S + i], q = Q[(size_t)l * S + i];
            if (pops) pops[(size_t)i * K + l] = mu;
            if (dpops) {
                if (mu == 0.0) { dpops[(size_t)i * K + l] = NAN; continue; }
                const double th_AA = q / (mu * mu), th_Al = q / mu, th_ll = theta[l * K + l];
                dpops[(size_t)i * K + l] = fabs(mu) * sqrt(fabs(th_AA + th_ll - 2.0 * th_Al));
            }
        }
    (void)f;
    return BCP_OK;
}

static int mbar_impl(const double* d_u, const int32_t* d_state, const int64_t* N_k, int K, int64_t N, int S, int device,
                     double tol, int max_iter, double* f_k, double* theta, double* pops, double* dpops, int32_t* iters,
                     cudaStream_t st) {
    MbarWork w;
    int rc = mbar_work_init(w, device, K, N, N_k, st);
    if (rc != BCP_OK) return rc;
    std::vector<double> th((size_t)K * K);
    int it = 0;
    rc = mbar_solve(w, d_u, N_k, tol, max_iter, f_k, th.data(), &it, st);
    if (iters) *iters = it;
    if (theta) std::copy(th.begin(), th.end(), theta);
    if (rc != BCP_OK) return rc;
    if (d_state && S > 0 && (pops || dpops)) {
        double *d_P = nullptr, *d_Q = nullptr;
        BCP_CUDA(bcp::pool_malloc((void**)&d_P, (size_t)K * S * 8, st));          // stream-ordered pool, no device sync
        cudaError_t e = bcp::pool_malloc((void**)&d_Q, (size_t)K * S * 8, st);
        if (e != cudaSuccess) { cudaFreeAsync(d_P, st); set_error("bcp_mbar: cudaMallocAsync: %s", cudaGetErrorString(e)); return BCP_ERR_CUDA; }
        rc = mbar_populations(w, d_u, f_k, d_state, S, d_P, d_Q, st);
        std::vector<double> P((size_t)K * S), Q((size_t)K * S);
        if (rc == BCP_OK) {
            e = cudaMemcpyAsync(P.data(), d_P, P.size() * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(Q.data(), d_Q, Q.size() * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error("bcp_mbar: copy back: %s", cudaGetErrorString(e)); rc = BCP_ERR_CUDA; }
        }
        cudaFreeAsync(d_P, st); cudaFreeAsync(d_Q, st);
        if (rc != BCP_OK) return rc;
        mbar_finish_outputs(K, S, f_k, th.data(), P.data(), Q.data(), pops, dpops);
    }
    return BCP_OK;
}

static int mbar_check(const void* u, const int64_t* N_k, int K, int64_t N, double* f_k) {
    BCP_REQUIRE(u && N_k && f_k, "bcp_mbar: NULL argument");
    BCP_REQUIRE(K >= 1 && K <= MB_MAXK, "bcp_mbar: K must be in 1..%d", MB_MAXK);
    int64_t tot = 0;
    for (int k = 0; k < K; ++k) { BCP_REQUIRE(N_k[k] >= 0, "bcp_mbar: N_k[%d] < 0", k); tot += N_k[k]; }
    BCP_REQUIRE(tot == N && N > 0, "bcp_mbar: sum(N_k) = %lld must equal N = %lld > 0", (long long)tot, (long long)N);
    BCP_REQUIRE(N_k[0] > 0, "bcp_mbar: N_k[0] must be > 0 (f_0 is the reference ensemble)");
    return BCP_OK;
}

extern "C" int bcp_mbar(const double* u_kn, const int64_t* N_k, int32_t K, int64_t N, const int32_t* state_n, int32_t S,
                        int device, double tol, int32_t max_iter, double* f_k, double* theta, double* pops,
                        double* dpops, int32_t* iters) {
    int rc = mbar_check(u_kn, N_k, K, N, f_k);
    if (rc != BCP_OK) return rc;
    if (state_n) for (int64_t n = 0; n < N; ++n) BCP_REQUIRE(state_n[n] >= 0 && state_n[n] < S, "bcp_mbar: state_n[%lld] out of range", (long long)n);
    DeviceGuard guard(device);
    double* d_u = nullptr; int32_t* d_s = nullptr;
    BCP_CUDA(dev_alloc(&d_u, (size_t)K * N));
    cudaError_t e = cudaMemcpy(d_u, u_kn, (size_t)K * N * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && state_n) { e = dev_alloc(&d_s, (size_t)N); if (e == cudaSuccess) e = cudaMemcpy(d_s, state_n, (size_t)N * 4, cudaMemcpyHostToDevice); }
    if (e != cudaSuccess) { cudaFree(d_u); cudaFree(d_s); set_error("bcp_mbar: upload: %s", cudaGetErrorString(e)); return BCP_ERR_CUDA; }
    rc = mbar_impl(d_u, d_s, N_k, K, N, S, device, tol > 0 ? tol : 1e-12, max_iter > 0 ? max_iter : 10000, f_k, theta, pops, dpops, iters, nullptr);
    cudaFree(d_u); cudaFree(d_s);
    return rc;
}

extern "C" int bcp_mbar_dev(const double* d_u_kn, const int64_t* N_k, int32_t K, int64_t N, const int32_t* d_state_n,
                            int32_t S, int device, void* stream, double tol, int32_t max_iter, double* f_k, double* theta,
                            double* pops, double* dpops, int32_t* iters) {
    int rc = mbar_check(d_u_kn, N_k, K, N, f_k);
    if (rc != BCP_OK) return rc;
    DeviceGuard guard(device);
    return mbar_impl(d_u_kn, d_state_n, N_k, K, N, S, device, tol > 0 ? tol : 1e-12, max_iter > 0 ? max_iter : 10000, f_k, theta,
                     pops, dpops, iters, (cudaStream_t)stream);
}

// one streaming reduction over a LOCAL shard of samples, for a multi-GPU solver that all-reduces
// s and M between passes: s[K] = sum_n W_nk, M[K][K] = sum_n W_nk W_nl (if want_hess), with
// a_k = f_k + ln N_k built from the GLOBAL N_k.
extern "C" int bcp_mbar_pass_dev(const double* d_u_kn, int64_t n_local, int32_t K, const int64_t* N_k_global,
                                 const double* f_k, int32_t want_hess, int device, void* stream, double* s_out,
                                 double* M_out) {
    BCP_REQUIRE(d_u_kn && N_k_global && f_k && s_out, "bcp_mbar_pass_dev: NULL argument");
    BCP_REQUIRE(K >= 1 && K <= MB_MAXK && n_local > 0, "bcp_mbar_pass_dev: bad K or n_local");
    DeviceGuard guard(device);
    MbarWork w;
    int rc = mbar_work_init(w, device, K, n_local, N_k_global, (cudaStream_t)stream);
    if (rc != BCP_OK) return rc;
    return mbar_pass(w, d_u_kn, f_k, want_hess != 0, s_out, M_out, (cudaStream_t)stream);
}

// the weighted state histograms of a LOCAL shard at free energies f: P[K][S] = sum_n W_nl [s_n = i] and
// Q[K][S] = sum_n W_nl^2 [s_n = i] (host outputs; a multi-GPU caller all-reduces both and finishes with
// P_i(l) = P, dP_i(l) = |P| sqrt(Q/P^2 + theta_ll - 2 Q/P))
extern "C" int bcp_mbar_pops_dev(const double* d_u_kn, int64_t n_local, int32_t K, const int64_t* N_k_global,
                                 const double* f_k, const int32_t* d_state_n, int32_t S, int device, void* stream,
                                 double* P_out, double* Q_out) {
    BCP_REQUIRE(d_u_kn && N_k_global && f_k && d_state_n && P_out && Q_out, "bcp_mbar_pops_dev: NULL argument");
    BCP_REQUIRE(K >= 1 && K <= MB_MAXK && n_local > 0 && S > 0, "bcp_mbar_pops_dev: bad K, n_local or S");
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    MbarWork w;
    int rc = mbar_work_init(w, device, K, n_local, N_k_global, st);
    if (rc != BCP_OK) return rc;
    double *d_P = nullptr, *d_Q = nullptr;
    BCP_CUDA(bcp::pool_malloc((void**)&d_P, (size_t)K * S * 8, st));
    cudaError_t e = bcp::pool_malloc((void**)&d_Q, (size_t)K * S * 8, st);
    if (e != cudaSuccess) { cudaFreeAsync(d_P, st); set_error("bcp_mbar_pops_dev: cudaMallocAsync: %s", cudaGetErrorString(e)); return BCP_ERR_CUDA; }
    rc = mbar_populations(w, d_u_kn, f_k, d_state_n, S, d_P, d_Q, st);
    if (rc == BCP_OK) {
        e = cudaMemcpyAsync(P_out, d_P, (size_t)K * S * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(Q_out, d_Q, (size_t)K * S * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("bcp_mbar_pops_dev: copy back: %s", cudaGetErrorString(e)); rc = BCP_ERR_CUDA; }
    }
    cudaFreeAsync(d_P, st); cudaFreeAsync(d_Q, st);
    return rc;
}
