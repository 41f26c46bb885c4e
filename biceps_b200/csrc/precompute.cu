// K1/K2/K4: per-state restraint-energy precompute.
//
// Replaces the Python loops of
//   Restraint_{cs,J,pf}.compute_sse           /root/reference/biceps/Restraint.py:344-353, 445-454, 736-741
//   Restraint_noe.compute_sse                 /root/reference/biceps/Restraint.py:552-569
//   build_exp_ref / compute_neglog_exp_ref    /root/reference/biceps/PosteriorSampler.py:74-104, Restraint.py:275-286
//   build_gaussian_ref / compute_neglog_...   /root/reference/biceps/PosteriorSampler.py:107-150, Restraint.py:288-298
//   compute_logZ                              /root/reference/biceps/PosteriorSampler.py:64-71
//
// All kernels are HBM-bound streaming reductions over model[S][J] (fp64, row-major):
//   row kernel    : L lanes per state row (L = 4..32 chosen from J), 16-byte double2 loads, shuffle reduction, one
//                   result (or G results for the gamma grid) per row, and in the SAME sweep the row's
//                   reference-potential sum;
//   column kernels: deterministic two-stage reduction (row-chunk partials, then a fixed-order combination); the
//                   Gaussian reference's mean and variance come out of ONE sweep (shifted sums per chunk, chunks
//                   combined with the pairwise-variance formula).
// An ensemble's restraint therefore costs at most two passes over model[S][J] (was 3 for the exponential and 4 for the
// Gaussian reference).
// The gamma-grid SSE uses the moment expansion  sum_j w (g e - m)^2 = g^2 Swee - 2 g Swem + Swmm
// (log-normal: d = ln(m/e), sum_j w (d - ln g)^2 = Swdd - 2 ln g Swd + ln^2 g Sw), so the kernel
// reads model once and writes sse[S][G] once instead of doing S*J*G fp64 operations.
#include "common.cuh"

#include <math.h>
#include <vector>

namespace bcp {

__device__ __forceinline__ double group_sum(double v, int L) {
    for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
    return v;
}

enum RowMode { ROW_SSE = 0, ROW_LIN = 1, ROW_GAMMA = 2, ROW_GAMMA_LOG = 3 };

// out[i] = c0 + sum_j a[j] * (model[i][j] - b[j])^2          (ROW_SSE)
// out[i] = c0 + sum_j a[j] * model[i][j]                     (ROW_LIN)
// out[i][g] = gam2[g]*Swee - 2*gam1[g]*Swem_i + Swmm_i       (ROW_GAMMA;   a = w, b = exp)
// out[i][g] = Swdd_i - 2*gam1[g]*Swd_i + gam2[g]*Sw          (ROW_GAMMA_LOG; gam1 = ln g, gam2 = ln^2 g)
// REF fuses the reference-potential sum of the same row into the sweep (one pass over model for both tables):
//   REF_LIN : ref_out[i] = rc0 + sum_j ra[j] * model[i][j]                   (exponential reference)
//   REF_SQ  : ref_out[i] = rc0 + sum_j ra[j] * (model[i][j] - rb[j])^2       (Gaussian reference)
enum RefMode { REF_NONE = 0, REF_LIN = 1, REF_SQ = 2 };
template <int MODE, int REF>
__global__ void __launch_bounds__(256)
row_kernel(const double* __restrict__ model, const double* __restrict__ a, const double* __restrict__ b,
           const double* __restrict__ c0_ptr, int S, int J, int L, const double* __restrict__ gam1,
           const double* __restrict__ gam2, int G, const double* __restrict__ see_ptr, double* __restrict__ out,
           const double* __restrict__ ra, const double* __restrict__ rb, const double* __restrict__ rc0_ptr,
           double* __restrict__ ref_out) {
    // scalars produced by the preceding tiny kernels stay on the device (no host round trip)
    const double c0 = c0_ptr ? *c0_ptr : 0.0;
    const double see = see_ptr ? *see_ptr : 0.0;
    const double rc0 = (REF != REF_NONE && rc0_ptr) ? *rc0_ptr : 0.0;
    const int lane = threadIdx.x & 31;
    const int sub = lane / L;                 // which row of this warp's pass
    const int l = lane - sub * L;
    const int rows_per_warp = 32 / L;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool vec2 = ((J & 1) == 0);         // rows stay 16-byte aligned only for even J
    for (long long base = (long long)warp * rows_per_warp; base < S; base += (long long)n_warps * rows_per_warp) {
        const long long i = base + sub;
        double s0 = 0.0, s1 = 0.0, sr = 0.0;
        if (i < S) {
            const double* row = model + (size_t)i * J;
            if (vec2) {
                const double2* row2 = reinterpret_cast<const double2*>(row);
                const double2* a2 = reinterpret_cast<const double2*>(a);
                const double2* b2 = reinterpret_cast<const double2*>(b);
                for (int j = l; j < (J >> 1); j += L) {
                    const double2 m = __ldcs(row2 + j);       // streamed once: evict-first
                    const double2 av = __ldg(a2 + j);
                    if (REF == REF_LIN) {
                        const double2 rv = __ldg(reinterpret_cast<const double2*>(ra) + j);
                        sr += rv.x * m.x + rv.y * m.y;
                    } else if (REF == REF_SQ) {
                        const double2 rv = __ldg(reinterpret_cast<const double2*>(ra) + j);
                        const double2 cv = __ldg(reinterpret_cast<const double2*>(rb) + j);
                        const double f0 = m.x - cv.x, f1 = m.y - cv.y;
                        sr += rv.x * f0 * f0 + rv.y * f1 * f1;
                    }
                    if (MODE == ROW_LIN) {
                        s0 += av.x * m.x + av.y * m.y;
                    } else {
                        const double2 bv = __ldg(b2 + j);
                        if (MODE == ROW_SSE) {
                            const double e0 = m.x - bv.x, e1 = m.y - bv.y;
                            s0 += av.x * e0 * e0 + av.y * e1 * e1;
                        } else if (MODE == ROW_GAMMA) {
                            s0 += av.x * m.x * m.x + av.y * m.y * m.y;
                            s1 += av.x * m.x * bv.x + av.y * m.y * bv.y;
                        } else {
                            const double d0 = log(m.x / bv.x), d1 = log(m.y / bv.y);
                            s0 += av.x * d0 * d0 + av.y * d1 * d1;
                            s1 += av.x * d0 + av.y * d1;
                        }
                    }
                }
            } else {
                for (int j = l; j < J; j += L) {
                    const double m = __ldcs(row + j);
                    const double av = __ldg(a + j);
                    if (REF == REF_LIN) sr += __ldg(ra + j) * m;
                    else if (REF == REF_SQ) { const double f0 = m - __ldg(rb + j); sr += __ldg(ra + j) * f0 * f0; }
                    if (MODE == ROW_LIN) {
                        s0 += av * m;
                    } else {
                        const double bv = __ldg(b + j);
                        if (MODE == ROW_SSE) {
                            const double e0 = m - bv;
                            s0 += av * e0 * e0;
                        } else if (MODE == ROW_GAMMA) {
                            s0 += av * m * m;
                            s1 += av * m * bv;
                        } else {
                            const double d0 = log(m / bv);
                            s0 += av * d0 * d0;
                            s1 += av * d0;
                        }
                    }
                }
            }
        }
        s0 = group_sum(s0, L);
        if (MODE >= ROW_GAMMA) s1 = group_sum(s1, L);
        if (REF != REF_NONE) {
            sr = group_sum(sr, L);
            if (i < S && l == 0) ref_out[i] = rc0 + sr;
        }
        if (i < S) {
            if (MODE == ROW_SSE || MODE == ROW_LIN) {
                if (l == 0) out[i] = c0 + s0;
            } else {
                double* o = out + (size_t)i * G;
                for (int g = l; g < G; g += L) {
                    const double g1 = __ldg(gam1 + g), g2 = __ldg(gam2 + g);
                    o[g] = (MODE == ROW_GAMMA) ? (g2 * see - 2.0 * g1 * s1 + s0) : (s0 - 2.0 * g1 * s1 + g2 * see);
                }
            }
        }
    }
}

// Column statistics in ONE sweep over model (stage 1: per row-chunk partials, thread = one column, coalesced over j):
//   psum[chunk][j] = sum of the chunk's rows;   VAR: pm2[chunk][j] = sum of squared deviations from the chunk's own
//   mean, accumulated against a shift (the chunk's first row) so that nothing cancels.
template <bool VAR>
__global__ void __launch_bounds__(256)
col_stats_kernel(const double* __restrict__ model, int S, int J, int rows_per_chunk, double* __restrict__ psum,
                 double* __restrict__ pm2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= J) return;
    const long long r0 = (long long)chunk * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < S ? r0 + rows_per_chunk : S;
    const double shift = VAR ? model[(size_t)r0 * J + j] : 0.0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
    long long i = r0;
    const double* __restrict__ col = model + j;
    for (; i + 7 < r1; i += 8) {                   // 8 independent loads in flight per thread (64 B; ~48 KB per SM)
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = col[(size_t)(i + k) * J];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] -= shift;
        a0 += v[0] + v[4]; a1 += v[1] + v[5]; a2 += v[2] + v[6]; a3 += v[3] + v[7];
        if (VAR) {
            q0 += v[0] * v[0] + v[4] * v[4]; q1 += v[1] * v[1] + v[5] * v[5];
            q2 += v[2] * v[2] + v[6] * v[6]; q3 += v[3] * v[3] + v[7] * v[7];
        }
    }
    for (; i < r1; ++i) {
        const double v = col[(size_t)i * J] - shift;
        a0 += v;
        if (VAR) q0 += v * v;
    }
    const double n = (double)(r1 - r0);
    const double sd = (a0 + a1) + (a2 + a3);       // sum of (x - shift)
    psum[(size_t)chunk * J + j] = n * shift + sd;
    if (VAR) pm2[(size_t)chunk * J + j] = ((q0 + q1) + (q2 + q3)) - sd * sd / n;
}

// stage 2: fixed-order combination of the chunks, then the per-observable reference-potential parameters.
//   kind 1 (exponential): p0[j] = beta_j = colsum / (S + 1)
//   kind 2 (gaussian)   : p0[j] = mean_j = colsum / S;  p1[j] = sqrt(M2_j / (S + 1)) with the chunks' M2 combined by
//                         M2 = sum_c M2_c + sum_c n_c (mean_c - mean)^2  (Chan et al.; PosteriorSampler.py:128-137)
__global__ void col_finish_kernel(const double* __restrict__ psum, const double* __restrict__ pm2, int n_chunks,
                                  int rows_per_chunk, int S, int J, int kind, double* __restrict__ p0, double* __restrict__ p1) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += psum[(size_t)c * J + j];
    if (kind == 1) { p0[j] = s / ((double)S + 1.0); return; }
    const double mean = s / (double)S;
    double m2 = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const long long r0 = (long long)c * rows_per_chunk;
        const double n = (double)((r0 + rows_per_chunk < S ? r0 + rows_per_chunk : S) - r0);
        const double dm = psum[(size_t)c * J + j] / n - mean;
        m2 += pm2[(size_t)c * J + j] + n * dm * dm;
    }
    p0[j] = mean;
    p1[j] = sqrt(m2 / ((double)S + 1.0));
}

// single small block: turns per-observable parameters into the row-kernel coefficient vectors.
//   exponential: coef[j] = w_j/beta_j, c0 = sum_j w_j ln beta_j
//   gaussian   : optional global sigma (mean_j sigma_j^-2)^-1/2, coef[j] = w_j/(2 sigma_j^2),
//                c0 = sum_j w_j (ln sqrt(2 pi) + ln sigma_j)
__global__ void ref_coef_kernel(int kind, int J, const double* __restrict__ w, double* __restrict__ p0,
                                double* __restrict__ p1, int use_global, double* __restrict__ coef,
                                double* __restrict__ c0_out) {
    __shared__ double red[256];
    const int t = threadIdx.x;
    if (kind == 2 && use_global) {
        double s = 0.0;
        for (int j = t; j < J; j += blockDim.x) s += 1.0 / (p1[j] * p1[j]);
        red[t] = s;
        __syncthreads();
        for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
        const double g = 1.0 / sqrt(red[0] / (double)J);
        __syncthreads();
        for (int j = t; j < J; j += blockDim.x) p1[j] = g;
        __syncthreads();
    }
    double c = 0.0;
    for (int j = t; j < J; j += blockDim.x) {
        if (kind == 1) {
            coef[j] = w[j] / p0[j];
            c += w[j] * log(p0[j]);
        } else {
            coef[j] = w[j] / (2.0 * p1[j] * p1[j]);
            c += w[j] * (0.91893853320467274178 + log(p1[j]));   // ln sqrt(2 pi)
        }
    }
    red[t] = c;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
    if (t == 0) *c0_out = red[0];
}

// see = sum_j w e^2 (normal) or sum_j w (log-normal); gam1/gam2 vectors for the expansion
__global__ void gamma_prep_kernel(int J, const double* __restrict__ w, const double* __restrict__ e,
                                  int log_normal, int G, const double* __restrict__ gamma,
                                  double* __restrict__ gam1, double* __restrict__ gam2, double* __restrict__ see) {
    __shared__ double red[256];
    const int t = threadIdx.x;
    double s = 0.0;
    for (int j = t; j < J; j += blockDim.x) s += log_normal ? w[j] : w[j] * e[j] * e[j];
    red[t] = s;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
    if (t == 0) *see = red[0];
    for (int g = t; g < G; g += blockDim.x) {
        const double x = log_normal ? log(gamma[g]) : gamma[g];
        gam1[g] = x;
        gam2[g] = x * x;
    }
}

// logZ[k] = ln sum_i exp(-energy[k][i])   (no max shift, like the reference; one block per lambda)
__global__ void logz_kernel(const double* __restrict__ energy, int S, double* __restrict__ logZ) {
    __shared__ double red[256];
    const double* e = energy + (size_t)blockIdx.x * S;
    double s = 0.0;
    for (int i = threadIdx.x; i < S; i += blockDim.x) s += exp(-e[i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) logZ[blockIdx.x] = log(red[0]);
}

static int lanes_per_row(int J) {
    const int work = (J & 1) ? J : J / 2;       // loads per row
    int L = 4;
    while (L < 32 && L * 2 <= work) L *= 2;     // at least ~2 loads per lane before widening
    return L;
}

// stream-ordered scratch (cudaMallocAsync pool): a plain cudaMalloc/cudaFree pair costs ~0.5 ms and an
// implicit device synchronisation per call, more than the kernels themselves
struct Scratch {
    std::vector<void*> p;
    cudaStream_t st = nullptr;
    ~Scratch() { for (void* q : p) cudaFreeAsync(q, st); }
    template <typename T> cudaError_t get(T** out, size_t n) {
        cudaError_t e = bcp::pool_malloc((void**)out, (n ? n : 1) * sizeof(T), st);
        if (e == cudaSuccess) p.push_back(*out);
        return e;
    }
};

// all pointers are device pointers; runs asynchronously on `st`
static int precompute_device(const bcp_obs& o, int S, int device, cudaStream_t st, double* d_sse, double* d_refsum,
                             double* d_par0, double* d_par1, Scratch& scr) {
    const int J = o.n_obs;
    const int sms = num_sms(device);
    const int L = lanes_per_row(J);
    const int rows_per_block = 8 * (32 / L);
    long long blocks = ((long long)S + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)sms * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
    int launches = 0;

    // reference potential first: column statistics in one sweep, then the coefficient vectors (tiny kernels)
    double *coef = nullptr, *c0 = nullptr;
    if (o.ref_kind != 0) {
        BCP_REQUIRE(d_refsum && d_par0 && (o.ref_kind == 1 || d_par1), "bcp_precompute: reference potential needs refsum/par outputs");
        // chunking: enough CTAs to fill the machine, chunks of >= 64 rows
        const int col_blocks = (J + 255) / 256;
        int n_chunks = (sms * 4 + col_blocks - 1) / col_blocks;
        if (n_chunks > (S + 63) / 64) n_chunks = (S + 63) / 64;
        if (n_chunks < 1) n_chunks = 1;
        const int rows_per_chunk = (S + n_chunks - 1) / n_chunks;
        n_chunks = (S + rows_per_chunk - 1) / rows_per_chunk;
        double *psum, *pm2 = nullptr;
        BCP_CUDA(scr.get(&psum, (size_t)n_chunks * J)); BCP_CUDA(scr.get(&coef, J)); BCP_CUDA(scr.get(&c0, 1));
        dim3 cgrid(col_blocks, n_chunks);
        if (o.ref_kind == 2) {
            BCP_CUDA(scr.get(&pm2, (size_t)n_chunks * J));
            col_stats_kernel<true><<<cgrid, 256, 0, st>>>(o.model, S, J, rows_per_chunk, psum, pm2);
        } else {
            col_stats_kernel<false><<<cgrid, 256, 0, st>>>(o.model, S, J, rows_per_chunk, psum, nullptr);
        }
        col_finish_kernel<<<col_blocks, 256, 0, st>>>(psum, pm2, n_chunks, rows_per_chunk, S, J, o.ref_kind, d_par0, d_par1);
        ref_coef_kernel<<<1, 256, 0, st>>>(o.ref_kind, J, o.weight, d_par0, d_par1, o.use_global_ref_sigma, coef, c0);
        launches += 3;
        BCP_CUDA(cudaGetLastError());
    }
    // one sweep over model for the SSE table AND the reference-potential sums
    double *g1 = nullptr, *g2 = nullptr, *see = nullptr;
    if (o.kind != BCP_KIND_SCALAR) {
        BCP_CUDA(scr.get(&g1, o.n_gamma)); BCP_CUDA(scr.get(&g2, o.n_gamma)); BCP_CUDA(scr.get(&see, 1));
        gamma_prep_kernel<<<1, 256, 0, st>>>(J, o.weight, o.exp, o.log_normal, o.n_gamma, o.gamma, g1, g2, see);
        ++launches;
    }
    const int G = o.kind == BCP_KIND_SCALAR ? 1 : o.n_gamma;
#define BCP_ROW(MODE, REF) row_kernel<MODE, REF><<<grid, 256, 0, st>>>(o.model, o.weight, o.exp, nullptr, S, J, L, g1, g2, G, see, d_sse, \
                                                                      coef, d_par0, c0, d_refsum)
#define BCP_ROW_REF(MODE) do { if (o.ref_kind == 0) BCP_ROW(MODE, REF_NONE); else if (o.ref_kind == 1) BCP_ROW(MODE, REF_LIN); \
                               else BCP_ROW(MODE, REF_SQ); } while (0)
    if (o.kind == BCP_KIND_SCALAR) BCP_ROW_REF(ROW_SSE);
    else if (o.log_normal) BCP_ROW_REF(ROW_GAMMA_LOG);
    else BCP_ROW_REF(ROW_GAMMA);
#undef BCP_ROW_REF
#undef BCP_ROW
    ++launches;
    BCP_CUDA(cudaGetLastError());
    BCP_LAUNCHED(launches);
    return BCP_OK;
}

}  // namespace bcp

using namespace bcp;

static int check_obs(const bcp_obs* o, int S) {
    BCP_REQUIRE(o && S > 0, "bcp_precompute: NULL observables or n_states <= 0");
    BCP_REQUIRE(o->kind == BCP_KIND_SCALAR || o->kind == BCP_KIND_GAMMA, "bcp_precompute: kind must be SCALAR or GAMMA");
    BCP_REQUIRE(o->n_obs > 0 && o->model && o->exp && o->weight, "bcp_precompute: bad n_obs / NULL pointer");
    BCP_REQUIRE(o->ref_kind >= 0 && o->ref_kind <= 2, "bcp_precompute: ref_kind must be 0, 1 or 2");
    if (o->kind == BCP_KIND_GAMMA) BCP_REQUIRE(o->n_gamma > 0 && o->gamma, "bcp_precompute: gamma grid missing");
    return BCP_OK;
}

extern "C" int bcp_precompute_dev(const bcp_obs* o, int32_t n_states, int device, void* stream, double* d_sse,
                                  double* d_refsum, double* d_ref_par0, double* d_ref_par1) {
    int rc = check_obs(o, n_states);
    if (rc != BCP_OK) return rc;
    BCP_REQUIRE(d_sse, "bcp_precompute_dev: NULL sse output");
    DeviceGuard guard(device);
    Scratch scr;
    scr.st = (cudaStream_t)stream;
    return precompute_device(*o, n_states, device, (cudaStream_t)stream, d_sse, d_refsum, d_ref_par0, d_ref_par1, scr);
}

extern "C" int bcp_precompute(const bcp_obs* o, int32_t n_states, int device, double* sse, double* refsum,
                              double* ref_par0, double* ref_par1) {
    int rc = check_obs(o, n_states);
    if (rc != BCP_OK) return rc;
    BCP_REQUIRE(sse, "bcp_precompute: NULL sse output");
    BCP_REQUIRE(o->ref_kind == 0 || (refsum && ref_par0 && (o->ref_kind == 1 || ref_par1)),
                "bcp_precompute: reference potential needs refsum / parameter outputs");
    DeviceGuard guard(device);
    Scratch scr;
    const int S = n_states, J = o->n_obs, G = o->kind == BCP_KIND_GAMMA ? o->n_gamma : 1;
    bcp_obs d = *o;
    double *dm, *de, *dw, *dg = nullptr, *dsse, *dref = nullptr, *dp0 = nullptr, *dp1 = nullptr;
    BCP_CUDA(scr.get(&dm, (size_t)S * J)); BCP_CUDA(scr.get(&de, J)); BCP_CUDA(scr.get(&dw, J));
    BCP_CUDA(cudaMemcpy(dm, o->model, (size_t)S * J * 8, cudaMemcpyHostToDevice));
    BCP_CUDA(cudaMemcpy(de, o->exp, (size_t)J * 8, cudaMemcpyHostToDevice));
    BCP_CUDA(cudaMemcpy(dw, o->weight, (size_t)J * 8, cudaMemcpyHostToDevice));
    if (o->kind == BCP_KIND_GAMMA) {
        BCP_CUDA(scr.get(&dg, G));
        BCP_CUDA(cudaMemcpy(dg, o->gamma, (size_t)G * 8, cudaMemcpyHostToDevice));
    }
    BCP_CUDA(scr.get(&dsse, (size_t)S * G));
    if (o->ref_kind) { BCP_CUDA(scr.get(&dref, S)); BCP_CUDA(scr.get(&dp0, J)); BCP_CUDA(scr.get(&dp1, J)); }
    d.model = dm; d.exp = de; d.weight = dw; d.gamma = dg;
    rc = precompute_device(d, S, device, nullptr, dsse, dref, dp0, dp1, scr);
    if (rc != BCP_OK) return rc;
    BCP_CUDA(cudaMemcpy(sse, dsse, (size_t)S * G * 8, cudaMemcpyDeviceToHost));
    if (o->ref_kind) {
        BCP_CUDA(cudaMemcpy(refsum, dref, (size_t)S * 8, cudaMemcpyDeviceToHost));
        BCP_CUDA(cudaMemcpy(ref_par0, dp0, (size_t)J * 8, cudaMemcpyDeviceToHost));
        if (o->ref_kind == 2) BCP_CUDA(cudaMemcpy(ref_par1, dp1, (size_t)J * 8, cudaMemcpyDeviceToHost));
    }
    return BCP_OK;
}

extern "C" int bcp_logz(const double* energy, int32_t n_lambda, int32_t n_states, int device, double* logZ) {
    BCP_REQUIRE(energy && logZ && n_lambda > 0 && n_states > 0, "bcp_logz: bad argument");
    DeviceGuard guard(device);
    Scratch scr;
    double *de, *dz;
    BCP_CUDA(scr.get(&de, (size_t)n_lambda * n_states)); BCP_CUDA(scr.get(&dz, n_lambda));
    BCP_CUDA(cudaMemcpy(de, energy, (size_t)n_lambda * n_states * 8, cudaMemcpyHostToDevice));
    logz_kernel<<<n_lambda, 256>>>(de, n_states, dz);
    BCP_LAUNCHED(1);
    BCP_CUDA(cudaGetLastError());
    BCP_CUDA(cudaMemcpy(logZ, dz, (size_t)n_lambda * 8, cudaMemcpyDeviceToHost));
    return BCP_OK;
}
