// Structures and device helpers shared by the generic and the specialised Metropolis kernels.
#pragma once
#include "common.cuh"
#include "tables.cuh"

namespace bcp {

struct ChainBuffers {
    long long n_chains;
    unsigned long long seed, chain_id0;
    const int* lam;         // [n]
    const int* group;       // [n]
    int* state;             // [n]
    int* isg;               // [R][n]
    int* igm;               // [R][n]
    int* ipf;               // [6][n]  PF grid indices (ensembles with a PFGRID restraint)
    double* E;              // [n]
    unsigned long long* accepted;     // [n]
    unsigned long long* sep;          // [R+1][n]  (this sample() call)
    unsigned long long* state_counts; // [groups][S]
    unsigned long long* param_counts; // [groups][HB]
    unsigned long long* spec_stats;   // [2]: batches run / batches rolled back by the speculative kernel (kernel choice)
};

struct LaunchArgs {
    unsigned long long step0;   // Philox step counter of local step 0
    long long s_begin, s_end;   // local step range of this launch
    long long burn;
    long long traj_every;       // 0 = none
    long long n_snap;
    double* traj_E;             // [n_snap][n]
    int* traj_state;            // [n_snap][n]
    unsigned char* traj_accept; // [n_snap][n]
    int* traj_idx;              // [n_snap][P][n]
    int* state_trace;           // [n_steps][trace_n] or nullptr
    long long trace_n;          // chains 0 .. trace_n-1 are traced
};

// number of recorded steps in [a, b): those with local step >= burn
__device__ __forceinline__ long long recorded(long long a, long long b, long long burn) {
    const long long lo = a > burn ? a : burn;
    return b > lo ? b - lo : 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 1-D TMA bulk copy global -> shared with mbarrier completion (SASS: UBLKCP).
__device__ __forceinline__ void stage_sigma_table(double* s_sig, const double* g_sig, int bytes,
                                                  uint64_t* mbar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(s_sig)), "l"(g_sig), "r"(bytes), "r"(smem_u32(mbar))
            : "memory");
    }
    // every thread waits for phase 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(mbar))
        : "memory");
}


// reference-stream mode (sampler_mt.cu): raw MT19937 outputs consumed by the single chain of the block
struct MtStream {
    const uint32_t* raw;    // [len] device
    long long len;
    long long* pos;         // device: outputs consumed so far (persists over launches)
    int* dry;               // device: set if the stream ran out
};
cudaError_t launch_mcmc_mt(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, const MtStream& M,
                           cudaStream_t st);

// launches the specialised kernel (sampler_fast.cu); returns cudaErrorNotSupported if the table
// layout is not the canonical one (at most one gamma restraint, and it is the last)
cudaError_t launch_mcmc_fast(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, int grid, int block,
                             cudaStream_t st);

// cooperative kernel (sampler_coop.cu): L lanes per chain; canonical layout only
cudaError_t launch_mcmc_coop(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms);

// speculative cooperative kernel (sampler_spec.cu): canonical layout, chain blocks in shared memory;
// cudaErrorNotSupported if the tables do not qualify (layout, shared-memory budget, division range)
cudaError_t launch_mcmc_spec(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms);
// warp-specialised speculative kernel (sampler_ws.cu): the spec kernel's critical path in a consumer warp, Philox /
// prefetch / histogram bookkeeping in a producer warp; same qualification rules as launch_mcmc_spec
cudaError_t launch_mcmc_ws(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st, int sms);
// table-driven kernel (sampler_tab.cu): one thread per chain on per-chain tables of the restraint terms, helper warps for
// everything else; canonical layouts with at most four scalar restraints, grids of at least 32 values
cudaError_t launch_mcmc_tab(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st);
// warp-per-chain kernel for ensembles with a PFGRID restraint (sampler_pf.cu); any layout
cudaError_t launch_mcmc_pf(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st);
// tuned variant for scalar restraints + one PFGRID restraint (sampler_pf2.cu); cudaErrorNotSupported otherwise
cudaError_t launch_mcmc_pf2(const SamplerTables& T, const ChainBuffers& C, const LaunchArgs& A, cudaStream_t st);
// -log P of explicit configurations when the ensemble has a PFGRID restraint (warp per configuration)
cudaError_t launch_energy_pf(const SamplerTables& T, long long N, const int* d_lam, const int* d_state,
                             const int* d_indices, double* d_out, int all_lambdas, cudaStream_t st);

}  // namespace bcp
