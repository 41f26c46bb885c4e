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

}  // namespace bcp

extern "C" int bcp_abi_version(void) { return BCP_ABI_VERSION; }
extern "C" const char* bcp_last_error(void) { return bcp::g_err; }
extern "C" uint64_t bcp_kernel_launches(void) { return bcp::g_kernel_launches.load(); }
