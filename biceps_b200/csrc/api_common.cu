// Error reporting and process-wide counters shared by all C-ABI translation units.
#include "common.cuh"

namespace bcp {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void keep_pool_memory(int device) {
    static std::atomic<uint64_t> done{0};
    if (device < 0 || device >= 64) return;
    const uint64_t bit = 1ull << device;
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    (void)cudaGetLastError();
    done.fetch_or(bit, std::memory_order_relaxed);
}

}  // namespace bcp

extern "C" int bcp_abi_version(void) { return BCP_ABI_VERSION; }
extern "C" const char* bcp_last_error(void) { return bcp::g_err; }
extern "C" uint64_t bcp_kernel_launches(void) { return bcp::g_kernel_launches.load(); }

// pinned (page-locked) host buffers for the callers' input / output arrays: host<->device copies of
// pageable memory run at a fraction of the PCIe rate
extern "C" void* bcp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        bcp::set_error("bcp_host_alloc: cudaHostAlloc(%zu) failed", bytes);
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void bcp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
