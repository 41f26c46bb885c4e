// K8/K9: convergence diagnostics of the sampled nuisance-parameter traces.
//
// Replaces the Python loops of /root/reference/biceps/convergence.py
//   g / compute_autocorrelation_curves / compute_autocorrelation_time      :78-125
//   compute_JSD and the fold x round bootstrap of Convergence.process      :145-186, :482-517
//
// K8 autocorrelation: C[tau] = sum_t z[t] z[t+tau] / (T - tau) for tau = 0 .. max_tau, z = f - mean(f), the
//   reference's quirk included (the LAST sample is left out of both factors, :108).  T * max_tau fp64 FMAs per
//   series: bound by the fp64 pipe, not by HBM (a series is read once per lag tile).  One CTA owns a tile of
//   1024 lags x 1024 time steps; a thread owns 8 consecutive lags and slides a register window over the
//   shared-memory copy of the series, so a step costs one broadcast load + one conflict-free load for 8 FMAs.
//   Partial sums per time segment are combined in a fixed order -> deterministic.
// K9 Jensen-Shannon divergence: one CTA per (parameter, fold, round) task builds the histograms of the whole
//   prefix and of its first part in shared memory and evaluates JSD = H - N1/N H1 - N2/N H2 with the
//   reference's term order.  The first part is the first half (:503-507), an explicit index list (the
//   reference's np.random.choice draw, made by the caller so that both use the same MT19937 stream), or a
//   Philox random half chosen on the device by a radix select over per-element keys.
#include "common.cuh"
#include "philox.cuh"

#include <math.h>
#include <stdlib.h>

namespace bcp {

// ------------------------------------------------------------------------------------------------
// K8
constexpr int AC_THREADS = 128;
constexpr int AC_TS = 1024;                        // time steps per CTA
// LPT lags per thread (8 or 16): a CTA covers AC_THREADS * LPT lags; the shifted factor's window is stored with
// one pad word per LPT entries, so the threads' stride is LPT + 1 doubles: conflict-free 64-bit loads
template <int LPT> struct AcCfg {
    static constexpr int LT = AC_THREADS * LPT;
    static constexpr int BW = AC_TS + LT;
    __host__ __device__ static constexpr int pad(int i) { return i + i / LPT; }
};

// mean (fixed-order block reduction) and centred copy z = f - mean of every series
__global__ void __launch_bounds__(256)
center_kernel(const double* __restrict__ f, long long T, double* __restrict__ z) {
    __shared__ double part[256];
    const double* s = f + (size_t)blockIdx.x * T;
    double* o = z + (size_t)blockIdx.x * T;
    double acc = 0.0;
    for (long long t = threadIdx.x; t < T; t += 256) acc += s[t];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    const double mean = part[0] / (double)T;
    for (long long t = threadIdx.x; t < T; t += 256) o[t] = s[t] - mean;
}

// partial[series][chunk][lag] = sum over the chunk's time segments of z[t] * z[t + lag], t + lag <= T - 2.
// A CTA walks seg_per_chunk consecutive segments of AC_TS time steps with its accumulators in registers.
template <int LPT>
__global__ void __launch_bounds__(AC_THREADS)
autocorr_partial_kernel(const double* __restrict__ z, long long T, int n_lags, int n_seg, int seg_per_chunk,
                        double* __restrict__ partial) {
    using Cfg = AcCfg<LPT>;
    __shared__ __align__(16) double A[AC_TS];
    __shared__ double B[Cfg::pad(Cfg::BW) + LPT];
    const int tile = blockIdx.x, chunk = blockIdx.y, series = blockIdx.z;
    const double* s = z + (size_t)series * T;
    const long long tau0 = (long long)tile * Cfg::LT;
    const long long last = T - 1;                   // index of the sample the reference leaves out
    const int j0 = threadIdx.x * LPT;               // first lag of this thread inside the tile
    double acc[LPT];
#pragma unroll
    for (int k = 0; k < LPT; ++k) acc[k] = 0.0;
    const int seg_end = min((chunk + 1) * seg_per_chunk, n_seg);
    for (int seg = chunk * seg_per_chunk; seg < seg_end; ++seg) {
        const long long t0 = (long long)seg * AC_TS;
        if (t0 + tau0 >= last) break;               // empty overlap from here on
        __syncthreads();
        for (int i = threadIdx.x; i < AC_TS; i += AC_THREADS) {
            const long long t = t0 + i;
            A[i] = (t < last) ? s[t] : 0.0;
        }
        for (int i = threadIdx.x; i < Cfg::BW; i += AC_THREADS) {
            const long long t = t0 + tau0 + i;
            B[Cfg::pad(i)] = (t < last) ? s[t] : 0.0;
        }
        __syncthreads();
        double w[2 * LPT];
#pragma unroll
        for (int k = 0; k < LPT; ++k) w[k] = B[Cfg::pad(j0 + k)];
#pragma unroll 2
        for (int t = 0; t < AC_TS; t += LPT) {
            // the next LPT window entries: indices j0 + t + LPT .. + 2 LPT - 1 share one pad offset
            const int base = Cfg::pad(j0 + t + LPT);
#pragma unroll
            for (int k = 0; k < LPT; ++k) w[LPT + k] = B[base + k];
#pragma unroll
            for (int tt = 0; tt < LPT; ++tt) {
                const double a = A[t + tt];
#pragma unroll
                for (int k = 0; k < LPT; ++k) acc[k] = fma(a, w[tt + k], acc[k]);
            }
#pragma unroll
            for (int k = 0; k < LPT; ++k) w[k] = w[LPT + k];
        }
    }
    double* out = partial + ((size_t)series * gridDim.y + chunk) * n_lags;
#pragma unroll
    for (int k = 0; k < LPT; ++k) {
        const long long lag = tau0 + j0 + k;
        if (lag < n_lags) out[lag] = acc[k];
    }
}

// curve[series][lag] = (sum over segments, in order) / (T - lag) [/ curve[0]]; tau[series] = sum of the curve
__global__ void __launch_bounds__(256)
autocorr_finish_kernel(const double* __restrict__ partial, long long T, int n_lags, int n_seg, int normalize,
                       double* __restrict__ curve, double* __restrict__ tau_sum) {
    const int series = blockIdx.x;
    const double* p = partial + (size_t)series * n_seg * n_lags;
    double* c = curve + (size_t)series * n_lags;
    __shared__ double c0;
    for (int lag = threadIdx.x; lag < n_lags; lag += 256) {
        double v = 0.0;
        for (int sgm = 0; sgm < n_seg; ++sgm) v += p[(size_t)sgm * n_lags + lag];
        v = v / (double)(T - lag);                  // 0/0 = nan at lag == T, -0.0 beyond, like numpy
        c[lag] = v;
        if (lag == 0) c0 = v;
    }
    __syncthreads();
    if (normalize) {
        const double d = c0;
        for (int lag = threadIdx.x; lag < n_lags; lag += 256) c[lag] = c[lag] / d;
        __syncthreads();
    }
    if (threadIdx.x == 0 && tau_sum) {              // Python's sum(): left to right (:123)
        double v = 0.0;
        for (int lag = 0; lag < n_lags; ++lag) v += c[lag];
        tau_sum[series] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K9
constexpr int JSD_THREADS = 256;
constexpr int JSD_MAX_BINS = 2048;

__device__ __forceinline__ uint64_t jsd_key(uint64_t seed, uint64_t stream, uint64_t e) {
    const Philox4 r = philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((uint64_t)r.w0 << 32) | (e & 0xffffffffull);
}

// sum_b -r/N ln(r/N) with nan -> 0, terms in bin order (convergence.py:177-182); all threads call it
__device__ double jsd_entropy(const int* hist, int nbins, double N, double* term) {
    for (int b = threadIdx.x; b < nbins; b += JSD_THREADS) {
        const double r = (double)hist[b];
        double h = 0.0;
        if (hist[b] > 0) h = __dmul_rn(__ddiv_rn(__dmul_rn(-1.0, r), N), log(__ddiv_rn(r, N)));
        term[b] = h;
    }
    __syncthreads();
    __shared__ double total;
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int b = 0; b < nbins; ++b) v = __dadd_rn(v, term[b]);
        total = v;
    }
    __syncthreads();
    const double out = total;
    __syncthreads();
    return out;
}

__global__ void __launch_bounds__(JSD_THREADS)
jsd_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ sel, const bcp_jsd_task* __restrict__ tasks,
           uint64_t seed, double* __restrict__ out, int32_t* __restrict__ bad) {
    __shared__ int h_tot[JSD_MAX_BINS];
    __shared__ int h_one[JSD_MAX_BINS];
    __shared__ double term[JSD_MAX_BINS];
    __shared__ unsigned int digit_count[256];
    __shared__ uint64_t s_prefix;
    __shared__ long long s_need;
    __shared__ int s_done;
    const bcp_jsd_task tk = tasks[blockIdx.x];
    const int nbins = tk.n_bins;
    const int32_t* x = idx + tk.offset;
    const long long n = tk.n_total, n1 = tk.n_first;
    for (int b = threadIdx.x; b < nbins; b += JSD_THREADS) {
        h_tot[b] = 0;
        h_one[b] = 0;
    }
    __syncthreads();
    int oob = 0;
    for (long long e = threadIdx.x; e < n; e += JSD_THREADS) {
        const int v = x[e];
        if (v < 0 || v >= nbins) { oob = 1; continue; }
        atomicAdd(&h_tot[v], 1);
        if (tk.mode == 0 && e < n1) atomicAdd(&h_one[v], 1);
    }
    if (tk.mode == 1) {
        const int32_t* sl = sel + tk.sel_offset;
        for (long long m = threadIdx.x; m < n1; m += JSD_THREADS) {
            const long long e = sl[m];
            if (e < 0 || e >= n) { oob = 1; continue; }
            const int v = x[e];
            if (v >= 0 && v < nbins) atomicAdd(&h_one[v], 1);
        }
    } else if (tk.mode == 2) {
        // radix select, most significant byte first: after the passes exactly n1 elements have
        // (key >> shift) < prefix
        if (threadIdx.x == 0) { s_prefix = 0; s_need = n1; s_done = (n1 <= 0 || n1 >= n); }
        __syncthreads();
        int shift = 64;
        if (!s_done) {
            for (int pass = 0; pass < 8; ++pass) {
                shift -= 8;
                for (int d = threadIdx.x; d < 256; d += JSD_THREADS) digit_count[d] = 0;
                __syncthreads();
                const uint64_t prefix = s_prefix;
                for (long long e = threadIdx.x; e < n; e += JSD_THREADS) {
                    const uint64_t k = jsd_key(seed, tk.stream, (uint64_t)e);
                    if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&digit_count[(k >> shift) & 0xff], 1u);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    long long need = s_need, cum = 0;
                    int d = 0;
                    for (; d < 255; ++d) {
                        if (cum + (long long)digit_count[d] > need) break;
                        cum += digit_count[d];
                    }
                    s_need = need - cum;
                    s_prefix = (prefix << 8) | (uint64_t)d;
                    s_done = (need - cum == 0);
                }
                __syncthreads();
                if (s_done) break;
            }
        }
        const uint64_t prefix = s_prefix;
        const bool all = (n1 >= n), none = (n1 <= 0);
        for (long long e = threadIdx.x; e < n; e += JSD_THREADS) {
            bool take = all;
            if (!all && !none) take = (jsd_key(seed, tk.stream, (uint64_t)e) >> shift) < prefix;
            const int v = x[e];
            if (take && v >= 0 && v < nbins) atomicAdd(&h_one[v], 1);
        }
    }
    if (oob) atomicExch(bad, 1);
    __syncthreads();
    const double N = (double)n, N1 = (double)n1, N2 = (double)(n - n1);
    const double H1 = jsd_entropy(h_one, nbins, N1, term);
    for (int b = threadIdx.x; b < nbins; b += JSD_THREADS) h_one[b] = h_tot[b] - h_one[b];
    __syncthreads();
    const double H2 = jsd_entropy(h_one, nbins, N2, term);
    const double H = jsd_entropy(h_tot, nbins, N, term);
    if (threadIdx.x == 0) {
        const double a = __dsub_rn(H, __dmul_rn(__ddiv_rn(N1, N), H1));
        out[blockIdx.x] = __dsub_rn(a, __dmul_rn(__ddiv_rn(N2, N), H2));
    }
}

// histogram of idx[off .. off+n) per task (block averaging, convergence.py:462-480)
__global__ void __launch_bounds__(JSD_THREADS)
hist_kernel(const int32_t* __restrict__ idx, const bcp_jsd_task* __restrict__ tasks, int64_t* __restrict__ out,
            int max_bins) {
    __shared__ int h[JSD_MAX_BINS];
    const bcp_jsd_task tk = tasks[blockIdx.x];
    for (int b = threadIdx.x; b < tk.n_bins; b += JSD_THREADS) h[b] = 0;
    __syncthreads();
    const int32_t* x = idx + tk.offset;
    for (long long e = threadIdx.x; e < tk.n_total; e += JSD_THREADS) {
        const int v = x[e];
        if (v >= 0 && v < tk.n_bins) atomicAdd(&h[v], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < max_bins; b += JSD_THREADS)
        out[(size_t)blockIdx.x * max_bins + b] = (b < tk.n_bins) ? h[b] : 0;
}

int autocorr_run(const double* d_series, int32_t n_series, int64_t T, int32_t max_tau, int32_t normalize,
                        cudaStream_t st, double* d_curves, double* d_tau) {
    const int n_lags = max_tau + 1;
    // 16 lags per thread once a CTA's 2048-lag tile is mostly used; BCP_AC_LPT overrides (tuning)
    int lpt = (n_lags > 1536) ? 16 : 8;
    if (const char* e = getenv("BCP_AC_LPT")) lpt = (atoi(e) == 16) ? 16 : 8;
    const int lt = AC_THREADS * lpt;
    const int n_tiles = (n_lags + lt - 1) / lt;
    const int n_seg = (int)((T + AC_TS - 1) / AC_TS);
    // time segments are grouped into chunks (one CTA each, accumulators stay in registers) so that the launch
    // still has a few waves of CTAs but the partial sums stay small: [n_series][n_chunk][n_lags]
    const long long want_ctas = 148LL * 5 * 6;
    long long n_chunk = (want_ctas + (long long)n_tiles * n_series - 1) / ((long long)n_tiles * n_series);
    if (n_chunk > n_seg) n_chunk = n_seg;
    if (n_chunk < 1) n_chunk = 1;
    const int seg_per_chunk = (int)((n_seg + n_chunk - 1) / n_chunk);
    n_chunk = (n_seg + seg_per_chunk - 1) / seg_per_chunk;
    double *z = nullptr, *partial = nullptr;
    BCP_CUDA(bcp::pool_malloc((void**)&z, sizeof(double) * (size_t)n_series * T, st));
    BCP_CUDA(bcp::pool_malloc((void**)&partial, sizeof(double) * (size_t)n_series * n_chunk * n_lags, st));
    center_kernel<<<n_series, 256, 0, st>>>(d_series, T, z);
    for (int s0 = 0; s0 < n_series; s0 += 65535) {          // gridDim.z limit
        const int ns = (n_series - s0 < 65535) ? n_series - s0 : 65535;
        dim3 grid(n_tiles, (unsigned)n_chunk, ns);
        if (lpt == 16)
            autocorr_partial_kernel<16><<<grid, AC_THREADS, 0, st>>>(
                z + (size_t)s0 * T, T, n_lags, n_seg, seg_per_chunk, partial + (size_t)s0 * n_chunk * n_lags);
        else
            autocorr_partial_kernel<8><<<grid, AC_THREADS, 0, st>>>(
                z + (size_t)s0 * T, T, n_lags, n_seg, seg_per_chunk, partial + (size_t)s0 * n_chunk * n_lags);
        BCP_LAUNCHED(1);
    }
    autocorr_finish_kernel<<<n_series, 256, 0, st>>>(partial, T, n_lags, (int)n_chunk, normalize, d_curves, d_tau);
    BCP_LAUNCHED(2);
    BCP_CUDA(cudaGetLastError());
    BCP_CUDA(cudaFreeAsync(z, st));
    BCP_CUDA(cudaFreeAsync(partial, st));
    return BCP_OK;
}

}  // namespace bcp

using namespace bcp;

extern "C" int bcp_autocorr_dev(const double* d_series, int32_t n_series, int64_t T, int32_t max_tau,
                                int32_t normalize, int device, void* stream, double* d_curves, double* d_tau) {
    BCP_REQUIRE(d_series && d_curves, "bcp_autocorr_dev: NULL pointer");
    BCP_REQUIRE(n_series > 0 && T >= 2 && max_tau >= 0, "bcp_autocorr_dev: need n_series > 0, T >= 2, max_tau >= 0");
        DeviceGuard g(device);
    return autocorr_run(d_series, n_series, T, max_tau, normalize, (cudaStream_t)stream, d_curves, d_tau);
}

extern "C" int bcp_autocorr(const double* series, int32_t n_series, int64_t T, int32_t max_tau, int32_t normalize,
                            int device, double* curves, double* tau) {
    BCP_REQUIRE(series && curves, "bcp_autocorr: NULL pointer");
    BCP_REQUIRE(n_series > 0 && T >= 2 && max_tau >= 0, "bcp_autocorr: need n_series > 0, T >= 2, max_tau >= 0");
        DeviceGuard g(device);
    cudaStream_t st = 0;
    const size_t n_in = (size_t)n_series * T, n_out = (size_t)n_series * (max_tau + 1);
    double *d_in = nullptr, *d_out = nullptr, *d_tau = nullptr;
    BCP_CUDA(pool_alloc(&d_in, n_in));
    BCP_CUDA(pool_alloc(&d_out, n_out));
    BCP_CUDA(pool_alloc(&d_tau, (size_t)n_series));
    BCP_CUDA(cudaMemcpyAsync(d_in, series, n_in * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = autocorr_run(d_in, n_series, T, max_tau, normalize, st, d_out, d_tau);
    if (rc == BCP_OK) {
        BCP_CUDA(cudaMemcpyAsync(curves, d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (tau) BCP_CUDA(cudaMemcpyAsync(tau, d_tau, (size_t)n_series * sizeof(double), cudaMemcpyDeviceToHost, st));
        BCP_CUDA(cudaStreamSynchronize(st));
    }
    pool_free(d_in);
    pool_free(d_out);
    pool_free(d_tau);
    return rc;
}

static int check_tasks(const bcp_jsd_task* tasks, int64_t n_tasks, int64_t n_idx, int64_t n_sel, const char* who) {
    for (int64_t i = 0; i < n_tasks; ++i) {
        const bcp_jsd_task& t = tasks[i];
        BCP_REQUIRE(t.n_bins > 0 && t.n_bins <= JSD_MAX_BINS, "%s: task %lld: n_bins must be in 1..%d", who,
                    (long long)i, JSD_MAX_BINS);
        BCP_REQUIRE(t.offset >= 0 && t.n_total >= 0 && t.offset + t.n_total <= n_idx,
                    "%s: task %lld reads outside idx", who, (long long)i);
        BCP_REQUIRE(t.n_first >= 0 && t.n_first <= t.n_total, "%s: task %lld: n_first out of range", who, (long long)i);
        BCP_REQUIRE(t.mode >= 0 && t.mode <= 2, "%s: task %lld: unknown mode %d", who, (long long)i, t.mode);
        BCP_REQUIRE(t.mode != 1 || (t.sel_offset >= 0 && t.sel_offset + t.n_first <= n_sel),
                    "%s: task %lld reads outside sel", who, (long long)i);
        BCP_REQUIRE(t.mode != 2 || t.n_total < (1LL << 32), "%s: task %lld: too long for the Philox split", who,
                    (long long)i);
    }
    return BCP_OK;
}

extern "C" int bcp_jsd(const int32_t* idx, int64_t n_idx, const int32_t* sel, int64_t n_sel,
                       const bcp_jsd_task* tasks, int64_t n_tasks, uint64_t seed, int device, double* jsd_out) {
    BCP_REQUIRE(idx && tasks && jsd_out && n_tasks > 0 && n_idx > 0, "bcp_jsd: NULL pointer or empty input");
    BCP_REQUIRE(n_sel == 0 || sel, "bcp_jsd: sel is NULL");
    int rc = check_tasks(tasks, n_tasks, n_idx, n_sel, "bcp_jsd");
    if (rc != BCP_OK) return rc;
    DeviceGuard g(device);
    cudaStream_t st = 0;
    int32_t *d_idx = nullptr, *d_sel = nullptr, *d_bad = nullptr;
    bcp_jsd_task* d_tasks = nullptr;
    double* d_out = nullptr;
    BCP_CUDA(pool_alloc(&d_idx, (size_t)n_idx));
    BCP_CUDA(pool_alloc(&d_sel, (size_t)(n_sel > 0 ? n_sel : 1)));
    BCP_CUDA(pool_alloc(&d_tasks, (size_t)n_tasks));
    BCP_CUDA(pool_alloc(&d_out, (size_t)n_tasks));
    BCP_CUDA(pool_alloc(&d_bad, (size_t)1));
    BCP_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    if (n_sel > 0) BCP_CUDA(cudaMemcpyAsync(d_sel, sel, sizeof(int32_t) * (size_t)n_sel, cudaMemcpyHostToDevice, st));
    BCP_CUDA(cudaMemcpyAsync(d_tasks, tasks, sizeof(bcp_jsd_task) * (size_t)n_tasks, cudaMemcpyHostToDevice, st));
    BCP_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), st));
    jsd_kernel<<<(unsigned)n_tasks, JSD_THREADS, 0, st>>>(d_idx, d_sel, d_tasks, seed, d_out, d_bad);
    BCP_LAUNCHED(1);
    BCP_CUDA(cudaGetLastError());
    int32_t bad = 0;
    BCP_CUDA(cudaMemcpyAsync(jsd_out, d_out, sizeof(double) * (size_t)n_tasks, cudaMemcpyDeviceToHost, st));
    BCP_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BCP_CUDA(cudaStreamSynchronize(st));
    pool_free(d_idx);
    pool_free(d_sel);
    pool_free(d_tasks);
    pool_free(d_out);
    pool_free(d_bad);
    BCP_REQUIRE(!bad, "bcp_jsd: a parameter index or a selection entry is out of range");
    return BCP_OK;
}

extern "C" int bcp_param_hist(const int32_t* idx, int64_t n_idx, const bcp_jsd_task* tasks, int64_t n_tasks,
                              int32_t max_bins, int device, int64_t* counts) {
    BCP_REQUIRE(idx && tasks && counts && n_tasks > 0 && n_idx > 0, "bcp_param_hist: NULL pointer or empty input");
    BCP_REQUIRE(max_bins > 0 && max_bins <= JSD_MAX_BINS, "bcp_param_hist: max_bins must be in 1..%d", JSD_MAX_BINS);
    int rc = check_tasks(tasks, n_tasks, n_idx, 0, "bcp_param_hist");
    if (rc != BCP_OK) return rc;
    for (int64_t i = 0; i < n_tasks; ++i)
        BCP_REQUIRE(tasks[i].n_bins <= max_bins && tasks[i].mode == 0, "bcp_param_hist: task %lld: n_bins > max_bins",
                    (long long)i);
    DeviceGuard g(device);
    cudaStream_t st = 0;
    int32_t* d_idx = nullptr;
    bcp_jsd_task* d_tasks = nullptr;
    int64_t* d_out = nullptr;
    BCP_CUDA(pool_alloc(&d_idx, (size_t)n_idx));
    BCP_CUDA(pool_alloc(&d_tasks, (size_t)n_tasks));
    BCP_CUDA(pool_alloc(&d_out, (size_t)n_tasks * max_bins));
    BCP_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    BCP_CUDA(cudaMemcpyAsync(d_tasks, tasks, sizeof(bcp_jsd_task) * (size_t)n_tasks, cudaMemcpyHostToDevice, st));
    hist_kernel<<<(unsigned)n_tasks, JSD_THREADS, 0, st>>>(d_idx, d_tasks, d_out, max_bins);
    BCP_LAUNCHED(1);
    BCP_CUDA(cudaGetLastError());
    BCP_CUDA(cudaMemcpyAsync(counts, d_out, sizeof(int64_t) * (size_t)n_tasks * max_bins, cudaMemcpyDeviceToHost, st));
    BCP_CUDA(cudaStreamSynchronize(st));
    pool_free(d_idx);
    pool_free(d_tasks);
    pool_free(d_out);
    return BCP_OK;
}
