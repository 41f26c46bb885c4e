// Latency / gather micro-benchmarks behind the K5 design:
// dependent-issue latency of the instructions on a Metropolis step's critical path, measured
// with clock64 on ONE warp, and whole-GPU random 8-byte gather throughput from an L2-resident
// table of the C3 size.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench
// scripts/microbench.cu ; run on the GPU box; prints one JSON object.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define N_DEP 2048

__device__ __forceinline__ double markstein_div(double a, double b, double y) {
    // y = RN(1/b) from a host table; two Newton corrections, the last one is Markstein's
    // correctly-rounding step (q faithful, y correctly rounded => RN(q + r*y) = RN(a/b))
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, y, q);
}

template <int OP>
__global__ void dep_kernel(double* out, long long* cycles, double seed, const double* tab, const int* chase, int n_iter) {
    __shared__ double s_tab[1024];
    __shared__ int s_chase[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { s_tab[i] = tab[i]; s_chase[i] = chase[i] & 1023; }
    __syncthreads();
    double x = seed + threadIdx.x * 1e-9;
    double y = 1.000000001;
    int k = threadIdx.x;
    float f = (float)seed;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n_iter; ++i) {
        if (OP == 0) x = __dadd_rn(x, y);
        if (OP == 1) x = __fma_rn(x, y, y);
        if (OP == 2) x = __dmul_rn(x, y);
        if (OP == 3) x = __ddiv_rn(x, y);
        if (OP == 4) x = markstein_div(x, y, 0.999999999);
        if (OP == 5) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 3, 4);
        if (OP == 6) k = s_chase[k];
        if (OP == 7) k = __ldg(chase + k);
        if (OP == 8) f = exp2f(f) * 0.5f;               // MUFU.EX2 + FMUL
        if (OP == 9) x = (x < 0.5) ? y : x;             // DSETP + 2 FSEL... (dependent select)
        if (OP == 10) k = (k * 3 + 1) & 0xffff;         // IMAD + LOP
        if (OP == 11) { float g = (float)x; x = (double)(g * 1.0000001f); }   // F2F both ways + FMUL
        if (OP == 12) x = s_tab[(int)(__double2loint(x) & 1023)] + 1e-300;   // LDS addressed by data + DADD
        if (OP == 13) { k = __shfl_sync(0xffffffffu, k, (threadIdx.x + 1) & 31); }
        if (OP == 14) { unsigned b = __ballot_sync(0xffffffffu, k & 1); k += __popc(b); }
        if (OP == 15) { k = __reduce_add_sync(0xffffffffu, k) & 0xffff; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    out[threadIdx.x] = x + k + f;
}

// whole-GPU random gathers: every thread walks its own LCG sequence over the table
__global__ void gather_kernel(const double* __restrict__ tab, uint32_t n_words, int per_thread, double* out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0.0;
#pragma unroll 8
    for (int i = 0; i < per_thread; ++i) {
        s = s * 1664525u + 1013904223u;
        acc += __ldg(tab + __umulhi(s, n_words));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// dependent gathers: the next address depends on the loaded value (one chain per thread)
__global__ void chase_kernel(const int* __restrict__ chase, int n, int hops, int* out, long long* cycles) {
    int k = (blockIdx.x * blockDim.x + threadIdx.x) % n;
    long long t0 = clock64();
    for (int i = 0; i < hops; ++i) k = __ldg(chase + k);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = k;
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int OP>
static double run_dep(double* d_out, long long* d_cyc, const double* d_tab, const int* d_chase, int threads) {
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
        dep_kernel<OP><<<1, threads>>>(d_out, d_cyc, 1.0, d_tab, d_chase, N_DEP);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        if (c < best) best = c;
    }
    return (double)best / N_DEP;
}

int main() {
    const int n_chase = 14 * 1024 * 1024 / 4;        // 14 MB of ints: the C3 table footprint
    std::vector<int> chase(n_chase);
    uint32_t s = 1;
    // random single-cycle permutation (Sattolo) so the chase never falls into a short loop
    for (int i = 0; i < n_chase; ++i) chase[i] = i;
    for (int i = n_chase - 1; i > 0; --i) {
        s = s * 1664525u + 1013904223u;
        int j = (int)(((uint64_t)s * (uint32_t)i) >> 32);
        int t = chase[i]; chase[i] = chase[j]; chase[j] = t;
    }
    std::vector<double> tab(n_chase / 2, 1.0);
    int* d_chase; double* d_tab; double* d_out; long long* d_cyc; int* d_iout;
    CK(cudaMalloc(&d_chase, n_chase * 4));
    CK(cudaMalloc(&d_tab, n_chase * 4));
    CK(cudaMalloc(&d_out, 1 << 24));
    CK(cudaMalloc(&d_iout, 1 << 24));
    CK(cudaMalloc(&d_cyc, 64));
    CK(cudaMemcpy(d_chase, chase.data(), n_chase * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tab, tab.data(), n_chase * 4, cudaMemcpyHostToDevice));

    printf("{\n \"dependent_issue_cycles_one_warp\": {\n");
    printf("  \"DADD\": %.2f,\n", run_dep<0>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"DFMA\": %.2f,\n", run_dep<1>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"DMUL\": %.2f,\n", run_dep<2>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"ddiv_rn\": %.2f,\n", run_dep<3>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"markstein_div_5op\": %.2f,\n", run_dep<4>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"SHFL_f64_width4\": %.2f,\n", run_dep<5>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"LDS_chase\": %.2f,\n", run_dep<6>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"LDG_chase_14MB_one_warp\": %.2f,\n", run_dep<7>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"MUFU_EX2_plus_FMUL\": %.2f,\n", run_dep<8>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"DSETP_select_f64\": %.2f,\n", run_dep<9>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"IMAD_LOP\": %.2f,\n", run_dep<10>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"F2F_f64_f32_f64_FMUL\": %.2f,\n", run_dep<11>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"LDS_by_data_plus_DADD\": %.2f,\n", run_dep<12>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"SHFL_b32\": %.2f,\n", run_dep<13>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"BALLOT_POPC\": %.2f,\n", run_dep<14>(d_out, d_cyc, d_tab, d_chase, 32));
    printf("  \"REDUX_ADD\": %.2f\n },\n", run_dep<15>(d_out, d_cyc, d_tab, d_chase, 32));

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    // random 8-byte gathers, table L2-resident (14 MB), grid sweep
    printf(" \"random_gather_8B_from_14MB_L2_table\": [\n");
    const int cfgs[][2] = {{148, 32}, {148, 128}, {592, 128}, {148 * 8, 256}, {148 * 16, 512}};
    for (int c = 0; c < 5; ++c) {
        const int grid = cfgs[c][0], block = cfgs[c][1], per = 4096;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            gather_kernel<<<grid, block>>>(d_tab, (uint32_t)(n_chase / 2), per, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double n = (double)grid * block * per;
        printf("  {\"grid\": %d, \"block\": %d, \"gathers_per_s\": %.4g, \"GBps_useful\": %.1f, \"GBps_sectors\": %.1f}%s\n",
               grid, block, n / (best * 1e-3), n * 8 / (best * 1e-3) / 1e9, n * 32 / (best * 1e-3) / 1e9, c < 4 ? "," : "");
    }
    printf(" ],\n");

    // dependent gathers: latency per hop at different occupancies
    printf(" \"dependent_gather_14MB\": [\n");
    const int ccfgs[][2] = {{1, 32}, {148, 32}, {148, 128}, {592, 256}, {148 * 8, 512}};
    for (int c = 0; c < 5; ++c) {
        const int grid = ccfgs[c][0], block = ccfgs[c][1], hops = 2048;
        float best = 1e30f; long long cyc = 0;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            chase_kernel<<<grid, block>>>(d_chase, n_chase, hops, d_iout, d_cyc);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) { best = ms; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); }
        }
        printf("  {\"grid\": %d, \"block\": %d, \"cycles_per_hop_warp0\": %.1f, \"hops_per_s\": %.4g}%s\n", grid, block,
               (double)cyc / hops, (double)grid * block * hops / (best * 1e-3), c < 4 ? "," : "");
    }
    printf(" ],\n \"sm_count\": %d, \"clock_khz_attr\": %d\n}\n", prop.multiProcessorCount, clk_khz);
    return 0;
}
