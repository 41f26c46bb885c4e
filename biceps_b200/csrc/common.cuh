// Shared host-side helpers for the biceps_b200 C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>

#include "../../include/biceps_b200.h"

namespace bcp {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_kernel_launches;

#define BCP_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            bcp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                           cudaGetErrorString(e__));                                      \
            return BCP_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define BCP_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            bcp::set_error(__VA_ARGS__);                                                  \
            return BCP_ERR_INVALID;                                                       \
        }                                                                                 \
    } while (0)

#define BCP_LAUNCHED(n) bcp::g_kernel_launches.fetch_add((n), std::memory_order_relaxed)

// RAII device-selection guard: every entry point runs on the handle's device and restores
// the caller's current device afterwards.

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
        (void)cudaGetLastError();   // a stale error of another library must not be blamed on our launches
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
static inline cudaError_t dev_alloc(T** p, size_t n) {
    *p = nullptr;
    if (n == 0) return cudaSuccess;
    return cudaMalloc((void**)p, n * sizeof(T));
}

// Stream-ordered allocation from the library's OWN memory pool (one per device, release threshold raised so that a
// create / destroy cycle recycles the same blocks and costs microseconds, where cudaMalloc / cudaFree cost ~0.5 ms each
// and synchronise the device).  The process-wide default pool of the device -- shared with torch's cudaMallocAsync
// backend, CuPy, RMM ... -- is left untouched.
cudaMemPool_t library_pool(int device);        // current device if device < 0; nullptr on failure (plain cudaMallocAsync is used)
cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t st);

template <typename T>
static inline cudaError_t pool_alloc(T** p, size_t n, cudaStream_t st = (cudaStream_t)0) {
    *p = nullptr;
    if (n == 0) return cudaSuccess;
    return pool_malloc((void**)p, n * sizeof(T), st);
}
static inline void pool_free(void* p, cudaStream_t st = (cudaStream_t)0) {
    if (p) cudaFreeAsync(p, st);
}

// K8 on device-resident series (convergence.cu): curves[n_series][max_tau+1], tau[n_series]
int autocorr_run(const double* d_series, int32_t n_series, int64_t T, int32_t max_tau, int32_t normalize,
                 cudaStream_t st, double* d_curves, double* d_tau);

static inline int num_sms(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n;
}

}  // namespace bcp
