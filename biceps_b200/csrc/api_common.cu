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

cudaMemPool_t library_pool(int device) {
    static std::atomic<cudaMemPool_t> pools[64];
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return nullptr;
    if (device < 0 || device >= 64) return nullptr;
    cudaMemPool_t p = pools[device].load(std::memory_order_acquire);
    if (p) return p;
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t fresh = nullptr;
    if (cudaMemPoolCreate(&fresh, &props) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(fresh, cudaMemPoolAttrReleaseThreshold, &thr);
    cudaMemPool_t expected = nullptr;
    if (!pools[device].compare_exchange_strong(expected, fresh, std::memory_order_acq_rel)) {
        cudaMemPoolDestroy(fresh);              // another thread was first
        return expected;
    }
    return fresh;
}

cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t st) {
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    cudaMemPool_t pool = library_pool(-1);
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, st) : cudaMallocAsync(p, bytes, st);
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
